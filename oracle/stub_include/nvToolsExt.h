/* stand-in for the NVTX header (absent from this image): the reference only brackets code with range markers */
#pragma once
static inline int nvtxRangePushA(const char *) { return 0; }
static inline int nvtxRangePop(void) { return 0; }
#define nvtxRangePush nvtxRangePushA
