#!/bin/sh
# Regenerates simple_tests_cases.json: the @cases array of the reference's own regression script
# (scripts/test/simple_tests.pl:66-4424: references, reads, arguments, expected hits) evaluated by perl and dumped as JSON.
# Data fixture only: tests/test_simple_tests_cases.py feeds these inputs to the reference program and to the policy engine.
set -e
REF=${1:-/root/reference}
{
  echo 'use strict; use warnings; use lib "'$REF'/scripts/test"; use DNA; use JSON::PP; use List::Util qw(max min);'
  sed -n '66,4424p' $REF/scripts/test/simple_tests.pl
  echo 'print JSON::PP->new->canonical->pretty->encode(\@cases);'
} > /tmp/_dump_cases.pl
perl /tmp/_dump_cases.pl > "$(dirname "$0")/simple_tests_cases.json"
