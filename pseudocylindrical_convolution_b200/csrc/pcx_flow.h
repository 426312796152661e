// pcx_flow.h - internal interface between pcx_ctx.cu (pcx_wave_decode) and pcx_flow.cu (the persistent dataflow decoder).
#pragma once
#include "pcx_common.cuh"

// Decodes every symbol of the nimg bitstreams with ONE persistent kernel (pcx_flow.cu).  *unsupported = true (and PCX_OK) when the
// shape does not fit this engine (too many images for the co-resident grid, shared memory): the caller falls back.
int pcx_wave_decode_flow(const pcx_wave_net &n, pcx_coder *const *coders, long long *n_symbols, cudaStream_t s, bool *unsupported);

// host coder threads for a call over nimg bitstreams (see pcx_flow.cu): shared by the encoder's and the step decoder's pool
int pcx_host_coder_threads(int nimg);
