// oracle/ref_glue_drv.cpp -- TEST INFRASTRUCTURE ONLY (see ref_glue.cpp).
// SwDriver::extend (aligner_sw_driver.cpp:299-484) exposed through the glue: how far a seed hit
// extends left (forward index) and right (mirror index) without an edit.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include <iostream>
#include <sstream>
#include <algorithm>
#include <limits>
#include <map>
#include <set>
#include <fstream>
#include <thread>
#include <mutex>
#include <atomic>
#include <array>
#include <utility>
#include <stdexcept>

#define private public
#define protected public
#include "aligner_sw_driver.h"
#undef private
#undef protected
#include "bt2_idx.h"
#include "reference.h"
#include "scoring.h"
#include "read.h"

struct RefHandleDrv {   // must mirror RefHandle in ref_glue.cpp
	std::unique_ptr<Ebwt> fw;
	std::unique_ptr<Ebwt> bw;
	std::unique_ptr<BitPairReference> ref;
	std::unique_ptr<Scoring> sc_e2e;
	std::unique_ptr<Scoring> sc_loc;
};

extern "C" {

void ref_extend(void* vh, const uint8_t* codes, int len, int fw, uint64_t off, uint64_t seedlen,
                uint64_t topf, uint64_t botf, uint64_t topb, uint64_t botb, uint64_t* nlex_nrex) {
	RefHandleDrv* h = (RefHandleDrv*)vh;
	static const char dna[] = "ACGTN";
	std::string s(len, 'N'), q(len, 'I');
	for(int i = 0; i < len; i++) s[i] = dna[codes[i] > 4 ? 4 : codes[i]];
	Read rd; rd.init("r", s.c_str(), q.c_str());
	SwDriver sd(1024 * 1024);
	PerReadMetrics prm;
	size_t nlex = 0, nrex = 0;
	sd.extend(rd, *h->fw, h->bw.get(), (TIndexOffU)topf, (TIndexOffU)botf, (TIndexOffU)topb, (TIndexOffU)botb,
	          fw != 0, (size_t)off, (size_t)seedlen, prm, nlex, nrex);
	nlex_nrex[0] = nlex; nlex_nrex[1] = nrex;
}

} // extern "C"
