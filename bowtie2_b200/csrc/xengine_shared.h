// xengine_shared.h -- host helpers shared by the two drivers of xengine.cuh (xengine_host.cpp, xengine.cu)
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/bt2g.h"

namespace xe {
struct XParams;
struct XTables { std::vector<int32_t> minsc, nceilRaw, ivalOne, ivalBoth; };
// bt2g_policy_params -> XParams with the SimpleFunc values (--score-min, --n-ceil, -i) tabulated per read length on the host
void buildParams(const bt2g_policy_params *pp, int offSize, int maxLen, XParams &P, XTables &T);
void scoringFromParams(const bt2g_policy_params *pp, bt2g_scoring *sc);
uint32_t genRandSeed(const uint8_t *codes, const uint8_t *quals, int len, const char *name, uint32_t seed);
}
