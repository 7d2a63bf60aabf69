"""INTEGRATION.md's whole-run example (the binding a bowtie2 maintainer would add around bt2g_stream_run) must compile against
include/bt2g.h as it stands."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_whole_run_example_compiles(tmp_path):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = [b for b in re.findall(r"```c\n(.*?)```", text, flags=re.S) if "bt2g_stream_run(" in b and "#include" in b]
    assert len(blocks) == 1
    src = tmp_path / "run.c"
    src.write_text(blocks[0])
    p = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
