// BSVD streaming engine (ring buffers + temporal-shift scatter stores) -- C ABI entry points.
#include "../../include/ss4k.h"

extern "C" {
int ss4k_bsvd_stream_open(ss4k_plan*, ss4k_bsvd_stream**) { return SS4K_E_INVALID; }
int ss4k_bsvd_stream_push(ss4k_bsvd_stream*, const void*, void*, int*, void*) { return SS4K_E_INVALID; }
int ss4k_bsvd_stream_flush(ss4k_bsvd_stream*, void*, int*, void*) { return SS4K_E_INVALID; }
int ss4k_bsvd_stream_reset(ss4k_bsvd_stream*) { return SS4K_E_INVALID; }
int ss4k_bsvd_stream_close(ss4k_bsvd_stream*) { return SS4K_E_INVALID; }
}
