// pcx_conv_tc.cu - pseudocylindrical convolution as an implicit GEMM on the 5th-generation tensor cores.
//
//   D[pixel, co] = sum over taps (ky,kx) and input channels ci of  X[y*s+ky, x*s+kx, ci] * W[co, ci, ky, kx]
//
// Activations are NHWC ("tile-major, channels last": [plane][row][column][channel]) so that a filter tap is a
// coordinate offset in the two OUTER dimensions of the TMA tensor map - TMA cannot start a box at an
// unaligned innermost coordinate (measured: illegal instruction), which rules out walking taps along a
// contiguous pixel axis.  Mapping (one CTA per SM, persistent over output tiles):
//   M = 128 output pixels  = bw columns x bh rows of one plane (bw*bh = 128, bw in {128, 64, 32})
//   N = NT output channels = 16 / 96 / 192 per tile (768-channel layers run as 4 N-tiles)
//   K = taps x Ci, streamed in blocks of 32 input channels of one tap
//   A operand: ONE TMA box (32 channels x bw x bh, SWIZZLE_128B) per K block straight from the halo-padded
//      activation buffer - 128 rows of 128 bytes, K-major; stride-2 layers use the tensor map's element strides.
//   B operand: weights repacked once to [tap][Co_pad][Ci] (K-major, SWIZZLE_128B), one TMA box NT x 32 per block.
//   D: fp32 accumulator in TMEM, 2 stages (2 x NT columns) so the epilogue of tile i overlaps the MMAs of i+1.
//   tcgen05.mma.cta_group::1.kind::tf32, M=128, N=NT, K=8: four per K block, issued by one thread.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2..5 epilogue (TMEM -> registers ->
// bias, PReLU / sigmoid, gate, residual, invalid-column zeroing -> 128-bit NHWC stores; lane = pixel).
#include "pcx_common.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace {

constexpr int BLOCK_M = 128;          // pixels per tile
constexpr int BLOCK_K = 32;           // input channels per pipeline stage (= one 128-byte swizzled row)
constexpr int UMMA_K = 8;             // tf32
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 4;     // 16 KB: 128 pixel rows x 128 B
constexpr int NUM_THREADS = 192;      // 6 warps (CTA-pair kernel: 4 epilogue warps per CTA, 8 per pair)
constexpr int TC_THREADS = 320;       // single-CTA kernel: producer, MMA issuer + EIGHT epilogue warps (two per TMEM lane quadrant)
constexpr int SQ_WARPS = 2;           // + two warps that square the activation stage in place (GDN: the GEMM reads x^2)
constexpr int MAX_CO_STAGED = 768;    // bias + PReLU slope vectors staged in shared memory (2 x 3 KB)
constexpr int VEC_SMEM = 2 * MAX_CO_STAGED * 4 + 4 * 32 * 36 * 4;   // + the epilogue warps' transpose tiles

struct TcParams {
    int planes, npart;
    int Ci, Co, Ho, Wo;
    int out_rows, out_pitch, out_y0, out_x0;
    int aux_rows, aux_pitch, aux_y0, aux_x0;
    int k, stride, act;
    int square;             // 1: the GEMM reads x * x (squarer warps, conv_tc_kernel<.., SQ = true>)
    int d2w;                // 1: write the result depth-to-space (Dtow stride 2 fused into the store), GEMM column q*Co/4 + c = channel 4c + q
    int bw, bh;             // tile = bw columns x bh rows, bw * bh = 128
    int tiles_x, tiles_y, n_tiles, co_pad;
    long long total_tiles;
    long long *dbg;         // optional per-CTA wait-time counters of the MMA thread (PCX_TC_DEBUG), else NULL
    int wl_out[PCX_MAX_PART];
};

template <int NT>
struct Cfg {
    static constexpr int B_STAGE_BYTES = NT * BLOCK_K * 4;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (184 * 1024) / STAGE_BYTES > 8 ? 8 : (184 * 1024) / STAGE_BYTES;
    static constexpr int TMEM_COLS = NT * 2 <= 32 ? 32 : (NT * 2 <= 64 ? 64 : (NT * 2 <= 128 ? 128 : (NT * 2 <= 256 ? 256 : 512)));
    static constexpr size_t SMEM = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + 256 /*barriers*/ + VEC_SMEM + 4 * 32 * 36 * 4 /*8 epilogue warps*/;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 operands, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) [4,6), a/b format TF32 (2) [7,10) [10,13),
// a_major K (0) [15], b_major K (0) [16], N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t instr_desc(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void stage_channel_vectors(float *s_bias, float *s_slope, const float *bias, const float *slope, int Co, int d2w)
{
    const int cq = Co >> 2;
    for (int i = threadIdx.x; i < MAX_CO_STAGED; i += blockDim.x) {
        const int src = (d2w && i < Co) ? 4 * (i % cq) + i / cq : i;       // GEMM column -> the layer's channel index
        s_bias[i] = (bias != nullptr && i < Co) ? bias[src] : 0.f;
        s_slope[i] = (slope != nullptr && i < Co) ? slope[src] : 1.f;
    }
}

struct Tile {
    long long plane;
    int y0, x0, n0;
};

__device__ __forceinline__ Tile decode_tile(const TcParams &p, long long t)
{
    Tile r;
    r.n0 = (int)(t % p.n_tiles); t /= p.n_tiles;
    r.x0 = (int)(t % p.tiles_x) * p.bw; t /= p.tiles_x;
    r.y0 = (int)(t % p.tiles_y) * p.bh; t /= p.tiles_y;
    r.plane = t;
    return r;
}

// a tile does tensor-core work only if it holds at least one valid pixel of its band
__device__ __forceinline__ bool tile_live(const TcParams &p, const Tile &t)
{
    return t.x0 < p.wl_out[(int)(t.plane % p.npart)];
}


// Epilogue of one warp = 32 consecutive pixels of one tile row (TMEM lanes) x NT channels of a finished accumulator.
// TMEM hands every LANE one pixel with 32 channels in registers; written like that, a warp-level 128-bit access touches 32
// different 768-byte-strided lines (the first version: the GDN / residual 1x1 layers ran at a quarter of the HBM rate).
// Here each 32-pixel x 32-channel chunk is transposed through a per-warp shared-memory tile (pitch 36 floats, conflict
// free both ways) so that 8 lanes cover the 128 contiguous bytes of one pixel: every global load (gate, residual) and
// store is four full 128-byte lines.  Math order per element is unchanged: +bias, act, *mul, +residual, fill.
// `bias` / `slope` point to the per-channel vectors staged in SHARED memory (broadcast LDS; the global copies would be
// re-fetched from L2 after every cluster-scope acquire, which invalidates L1).
// Transcendental epilogues on the SFU (MUFU.RSQ / SQRT / EX2 / RCP, ~1e-7 relative error): the operands of these layers
// went through TF32 tensor-core products (1e-3), and the IEEE-rounded 1/sqrtf / expf / division sequences (~43 instructions
// per channel) made the GDN epilogue issue-bound - as expensive as the HBM traffic of the layer (SASS count, r1y).
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float r;
    asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// The epilogue's transpose tile is addressed through explicit shared-state-space instructions: the kernels derive their
// shared-memory pointers from an integer-aligned base, so plain dereferences compile to GENERIC LD / ST, which the compiler
// must keep ordered against the global stores of the previous pixel - the per-pixel LDS -> SFU -> STG chains ran one
// after the other (SASS of r2a) and the few warps of the CTA could not hide their latency.
__device__ __forceinline__ float4 lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w));
}

constexpr int EPI_PITCH = 36;
constexpr int EPI_SMEM = 4 * 32 * EPI_PITCH * 4;       // four epilogue warps

template <int NT, int ACTK, bool DBG = false, bool PIPE = false>
__device__ __forceinline__ void epilogue_warp(const TcParams &p, long long plane, int oy, int ox0, int n0, bool live, int wl, uint32_t taddr0,
                                              float *__restrict__ stage, const float *__restrict__ bias, const float *__restrict__ slope,
                                              const float *__restrict__ mul, const float *__restrict__ residual, float *__restrict__ y,
                                              long long *t_ld = nullptr, int step_first = 0, int step_stride = 1)
{
    const int lane = threadIdx.x & 31;
    const int nco = min(NT, p.Co - n0 * NT);       // multiple of 4
    const int cbase = n0 * NT;
    const bool row_ok = oy < p.Ho;
    // lane -> (pixel sub-index, 16-byte channel chunk) of the transposed view
    const int psub = lane >> 3, cch = lane & 7;
    // depth-to-space store (Dtow, dtow_cuda.cu:38-55, fused): the N tile n0 holds the channels 4c + n0 of the layer, i.e. output
    // pixel (2 oy + n0 / 2, 2 ox + n0 % 2), channel c of a plane of Co / 4 channels - still whole 128-byte lines per store
    const int ostride = p.d2w ? p.Co >> 1 : p.Co;          // floats between the stores of consecutive tile pixels
    float *yrow = p.d2w ? y + (((plane * p.out_rows + 2 * oy + (n0 >> 1) + p.out_y0) * (long long)p.out_pitch) + 2 * ox0 + (n0 & 1) + p.out_x0) * (p.Co >> 2) + cch * 4
                        : y + (((plane * p.out_rows + oy + p.out_y0) * (long long)p.out_pitch) + ox0 + p.out_x0) * p.Co + cbase + cch * 4;
    const long long arow = (((plane * p.aux_rows + oy + p.aux_y0) * (long long)p.aux_pitch) + ox0 + p.aux_x0) * p.Co + cbase + cch * 4;
    constexpr int STEP = NT >= 32 ? 32 : 16;
    constexpr int CCH = STEP / 4;                  // 16-byte chunks per pixel and step (8 or 4)
    // the gate operand only ever accompanies a transcendental epilogue (sigmoid gate of the attention block, GDN / IGDN):
    // the plain / PReLU instances do not carry its registers
    constexpr bool CAN_MUL = ACTK != 0;
    const bool has_mul = CAN_MUL && mul != nullptr;
    // A layer whose epilogue reads a gate and / or a residual (GDN, attention gate, residual 1x1) is bound by these loads,
    // not by the tensor pipe: with four warps issuing them AFTER the TMEM read the 1x1 layers ran at ~0.4 of the HBM floor
    // and the warps stalled on the long scoreboard (launch list profiles/r1y_*, ncu profiles/r1z_conv_tc_gdn.json).
    //   * the single-CTA kernel runs two warps per lane quadrant, each taking every other 32-channel step
    //     (step_first / step_stride);
    //   * the operand loads are software-pipelined by HALF steps: a lane's eight pixels form two groups of four; the next
    //     step's loads of a group are issued as soon as the current step has consumed that group's registers, so 8..16
    //     independent 128-bit loads per lane are in flight through the TMEM read, the transpose and the arithmetic.
    const uint32_t stage_s = smem_u32(stage);
    const int xlim = min(wl, p.Wo) - ox0 - psub;          // pixel i of this lane is valid iff 4 i < xlim
    const int istep = 4 * p.Co;
    const bool any_aux = has_mul || residual != nullptr;
    float4 m4[CAN_MUL ? 8 : 1], a4[8];
    auto load_aux = [&](int c0, int grp) {
        const int co = c0 + cch * 4;
        if (any_aux && row_ok && cch < CCH && co < nco) {
            const long long ao = arow + (long long)psub * p.Co + c0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = grp * 4 + k;
                const bool ok = live && 4 * i < xlim;
                if (CAN_MUL) m4[i] = (has_mul && ok) ? __ldg(reinterpret_cast<const float4 *>(mul + ao + i * istep)) : make_float4(1.f, 1.f, 1.f, 1.f);
                a4[i] = (residual && ok) ? __ldg(reinterpret_cast<const float4 *>(residual + ao + i * istep)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    const int c_first = step_first * STEP, c_inc = step_stride * STEP;
    if (PIPE && c_first < NT) { load_aux(c_first, 0); load_aux(c_first, 1); }
#pragma unroll 1
    for (int c0 = c_first; c0 < NT; c0 += c_inc) {
        const int co = c0 + cch * 4;
        const bool mine = row_ok && cch < CCH && co < nco;
        if (!PIPE) { load_aux(c0, 0); load_aux(c0, 1); }     // CTA-pair kernel: tensor-bound, the plain order is the leaner one
        if (live) {
            uint32_t v[STEP];
            const uint32_t taddr = taddr0 + (uint32_t)c0;
            long long t0 = 0;
            if (DBG) t0 = clock64();
            if constexpr (STEP == 32) tmem_ld32(taddr, v);
            else tmem_ld16(taddr, v);
            tmem_wait_ld();
            if (DBG) *t_ld += clock64() - t0;
            const uint32_t sw = stage_s + (uint32_t)(lane * EPI_PITCH * 4);
#pragma unroll
            for (int j = 0; j < STEP; j += 4)
                sts_f4(sw + j * 4, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        __syncwarp();
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = b4;
        if (mine) {
            if (bias) b4 = *reinterpret_cast<const float4 *>(bias + cbase + co);
            if (p.act == 1) s4 = *reinterpret_cast<const float4 *>(slope + cbase + co);
        }
#pragma unroll
        for (int grp = 0; grp < 2; grp++) {
            if (mine) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int i = grp * 4 + k;
                    const int px = i * 4 + psub;
                    const int ox = ox0 + px;
                    // straight-line arithmetic on every pixel (invalid ones are zeroed by a select, out-of-row ones skip the
                    // store): with branches around each pixel the compiler serialised the four LDS -> SFU -> STG chains and
                    // the ten warps of the CTA could not hide their latency
                    float4 r = lds_f4(stage_s + (uint32_t)((px * EPI_PITCH + cch * 4) * 4));
                    if (bias) { r.x = __fadd_rn(r.x, b4.x); r.y = __fadd_rn(r.y, b4.y); r.z = __fadd_rn(r.z, b4.z); r.w = __fadd_rn(r.w, b4.w); }
                    if (p.act == 1) {
                        r.x = r.x < 0.f ? __fmul_rn(r.x, s4.x) : r.x;
                        r.y = r.y < 0.f ? __fmul_rn(r.y, s4.y) : r.y;
                        r.z = r.z < 0.f ? __fmul_rn(r.z, s4.z) : r.z;
                        r.w = r.w < 0.f ? __fmul_rn(r.w, s4.w) : r.w;
                    } else if (ACTK == 2) {     // sigmoid / rsqrt / sqrt live in separate kernel instances (ACTK = the act code):
                        r.x = fast_sigmoid(r.x); r.y = fast_sigmoid(r.y);      // inlined together they pushed the kernels out of
                        r.z = fast_sigmoid(r.z); r.w = fast_sigmoid(r.w);      // the instruction cache
                    } else if (ACTK == 3) {
                        r.x = fast_rsqrt(r.x); r.y = fast_rsqrt(r.y); r.z = fast_rsqrt(r.z); r.w = fast_rsqrt(r.w);
                    } else if (ACTK == 4) {
                        r.x = fast_sqrt(r.x); r.y = fast_sqrt(r.y); r.z = fast_sqrt(r.z); r.w = fast_sqrt(r.w);
                    }
                    if (CAN_MUL && has_mul) { r.x = __fmul_rn(r.x, m4[i].x); r.y = __fmul_rn(r.y, m4[i].y); r.z = __fmul_rn(r.z, m4[i].z); r.w = __fmul_rn(r.w, m4[i].w); }
                    if (residual) { r.x = __fadd_rn(a4[i].x, r.x); r.y = __fadd_rn(a4[i].y, r.y); r.z = __fadd_rn(a4[i].z, r.z); r.w = __fadd_rn(a4[i].w, r.w); }
                    if (!(live && ox < wl)) r = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ox < p.Wo) *reinterpret_cast<float4 *>(yrow + (long long)px * ostride + c0) = r;
                }
            }
            // this group's operand registers are free: the next step's loads go out now
            if (PIPE && c0 + c_inc < NT) load_aux(c0 + c_inc, grp);
        }
        __syncwarp();
    }
}

// SQ: the GEMM runs on x * x (GDN / IGDN, PseudoContextV2.py:186-216: beta' + gamma' x^2).  The round-1 form materialised x^2 in
// HBM with its own kernel (one read + one write of the activation, then the GEMM read it back); here two extra warps square
// each 16 KB activation stage IN SHARED MEMORY between the TMA's arrival and the MMA: same fp32 product, same tf32 operand,
// no HBM traffic (two warps: 384 threads keep the 168 registers of the epilogue).  Pipeline per stage: TMA -> full_bar -> squarer warps (generic-proxy read-modify-write, fence.proxy.async) ->
// sq_bar -> tcgen05.mma -> empty_bar.
template <int NT, int ACTK, bool SQ = false>
__global__ void __launch_bounds__(TC_THREADS + (SQ ? SQ_WARPS * 32 : 0), 1) conv_tc_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                  const __grid_constant__ CUtensorMap map_w, TcParams p,
                                                                  const float *__restrict__ bias, const float *__restrict__ slope,
                                                                  const float *__restrict__ mul, const float *__restrict__ residual,
                                                                  float *__restrict__ y)
{
    using C = Cfg<NT>;
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operands need 1024-byte alignment
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_base = smem;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)C::STAGES * C::STAGE_BYTES);
    uint64_t *empty_bar = full_bar + C::STAGES;
    uint64_t *acc_full = empty_bar + C::STAGES;     // [2]
    uint64_t *acc_empty = acc_full + 2;             // [2]
    uint64_t *sq_bar = acc_empty + 2;               // [STAGES] (SQ only): the stage's activations have been squared
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sq_bar + C::STAGES);
    static_assert((3 * C::STAGES + 4) * 8 + 4 <= 256, "barrier block");
    float *s_bias = reinterpret_cast<float *>(smem + (size_t)C::STAGES * C::STAGE_BYTES + 256);
    float *s_slope = s_bias + MAX_CO_STAGED;
    float *s_stage = s_slope + MAX_CO_STAGED;
    stage_channel_vectors(s_bias, s_slope, bias, slope, p.Co, p.d2w);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_w);
        for (int s = 0; s < C::STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&sq_bar[s], SQ_WARPS);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 8);           // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int taps = p.k * p.k;
    const int kblocks = p.Ci / BLOCK_K;
    const int iters = taps * kblocks;

    if (warp == 0) {
        // ===================================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                Tile tl = decode_tile(p, t);
                if (!tile_live(p, tl)) continue;
                int tap = 0, kb = 0, ky = 0, kx = 0;
                for (int it = 0; it < iters; it++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char *sa = stage_base + (size_t)stage * C::STAGE_BYTES;
                    unsigned char *sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    tma_load_4d(sa, &map_x, &full_bar[stage], kb * BLOCK_K, tl.x0 * p.stride + kx, tl.y0 * p.stride + ky, (int)tl.plane);
                    tma_load_3d(sb, &map_w, &full_bar[stage], kb * BLOCK_K, tl.n0 * NT, tap);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    if (++kb == kblocks) {
                        kb = 0; ++tap;
                        if (++kx == p.k) { kx = 0; ++ky; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================================== MMA issuer
        // one thread runs the whole loop; descriptors are `stage base + constant` (see conv_pair_kernel)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            constexpr uint32_t idesc = instr_desc(NT);
            const uint64_t desc_hi = smem_desc(0, 16, 1024) & 0xFFFFFFFF00000000ull;
            const uint32_t lo0 = (smem_u32(stage_base) >> 4) & 0x3fff;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                Tile tl = decode_tile(p, t);
                if (!tile_live(p, tl)) continue;
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);        // epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * NT;
                for (int it = 0; it < iters; it++) {
                    mbar_wait(SQ ? &sq_bar[stage] : &full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = lo0 + (uint32_t)stage * (C::STAGE_BYTES >> 4);
                    const uint32_t b_lo = a_lo + (A_STAGE_BYTES >> 4);
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / UMMA_K; kk++) {
                        // both operands K-major SWIZZLE_128B: 8 tf32 = 32 B further inside the 128 B row per MMA,
                        // 8-row groups (one swizzle atom) 1024 B apart
                        const uint64_t ad = desc_hi | (uint64_t)(a_lo + (uint32_t)((kk * UMMA_K * 4) >> 4));
                        const uint64_t bd = desc_hi | (uint64_t)(b_lo + (uint32_t)((kk * UMMA_K * 4) >> 4));
                        umma_tf32(tmem_d, ad, bd, idesc, (it | kk) != 0);
                    }
                    umma_commit(&empty_bar[stage]);               // smem slot free once these MMAs are done
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[acc]);                      // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (SQ && warp >= TC_THREADS / 32) {
        // ===================================================================================== squarer warps
        const int tq = threadIdx.x - TC_THREADS;
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            Tile tl = decode_tile(p, t);
            if (!tile_live(p, tl)) continue;
            for (int it = 0; it < iters; it++) {
                mbar_wait(&full_bar[stage], phase);
                float4 *a = reinterpret_cast<float4 *>(stage_base + (size_t)stage * C::STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < A_STAGE_BYTES / 16 / (SQ_WARPS * 32); i++) {
                    float4 v = a[tq + i * SQ_WARPS * 32];
                    v.x = __fmul_rn(v.x, v.x); v.y = __fmul_rn(v.y, v.y); v.z = __fmul_rn(v.z, v.z); v.w = __fmul_rn(v.w, v.w);
                    a[tq + i * SQ_WARPS * 32] = v;
                }
                fence_proxy_async();                              // generic-proxy stores -> the MMA's async-proxy operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&sq_bar[stage]);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================================================== epilogue warps
        const int q = warp & 3;                 // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;       // warps 2..5 take the even 32-channel steps of a tile, warps 6..9 the odd ones
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            Tile tl = decode_tile(p, t);
            const int g = (int)(tl.plane % p.npart);
            const int wl = p.wl_out[g];
            const int m0 = q * 32;                       // first pixel of this warp: 32 consecutive pixels of one tile row (32 | bw)
            const int oy = tl.y0 + m0 / p.bw;
            const int ox0 = tl.x0 + m0 % p.bw;
            const bool live = tile_live(p, tl);
            if (live) {
                mbar_wait(&acc_full[acc], acc_phase);
                tc_fence_after();
            }
            epilogue_warp<NT, ACTK, false, true>(p, tl.plane, oy, ox0, tl.n0, live, wl, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT),
                                     s_stage + (warp - 2) * 32 * EPI_PITCH, bias ? s_bias : nullptr, s_slope, mul, residual, y, nullptr, half, 2);
            if (live) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}


// ================================================================================================ 3x3 stride-1, CTA pairs
// conv_pair_kernel: the 3x3 / stride-1 layers (the bulk of the transforms' FLOPs) on `cta_group::2` MMAs.
//
// The single-CTA kernel above is bound by L2 -> shared-memory bandwidth, not by the tensor pipe (ncu, profiles/r1a_*:
// tensor pipe 44 % active, 2.2 MB of operand traffic per 128-pixel tile against ~43 B/clk/SM of L2 bandwidth).  Two
// changes cut the traffic 2.3x:
//   * a CTA pair (cluster of 2, one tile row each) shares every weight block: each CTA stages only HALF of the
//     N = NT weight rows and `tcgen05.mma.cta_group::2` (M = 256) reads both halves;
//   * the activation tile is staged ONCE per 32-channel block as a halo tile [3 rows][130 pixels][32 ch] (one TMA
//     box, SWIZZLE_128B) and the nine filter taps are nine shared-memory DESCRIPTORS into it (start address shifted
//     by (ky*130 + kx) pixel rows of 128 B) instead of nine TMA loads.
// Pipelines: A halo ring (2 stages), B ring (8 stages of NT/2 x 32), TMEM accumulator ring (2 stages) - all mbarrier
// based; full barriers live in the leader CTA (rank 0), whose warp 1 issues the MMAs for the pair; empty / accumulator
// barriers are signalled in both CTAs by multicast commits.
constexpr int HALO_W = BLOCK_M + 2;                         // pixels per halo-tile row
constexpr int A2_BYTES = 3 * HALO_W * BLOCK_K * 4;          // 49,920 B landed by one TMA box
constexpr int A2_STAGE = (A2_BYTES + 1023) / 1024 * 1024;   // 50,176 B (swizzle atoms stay 1024-aligned)
constexpr int A2_STAGES = 2;
constexpr int B2_STAGES = 8;

template <int NT>
struct Cfg2 {
    static constexpr int B_HALF = NT / 2 * BLOCK_K * 4;     // bytes of weights staged per CTA and (tap, k-block)
    static constexpr int B_STAGE = B_HALF < 1024 ? 1024 : B_HALF;
    static constexpr int TMEM_COLS = Cfg<NT>::TMEM_COLS;
    static constexpr size_t SMEM = 1024 + (size_t)A2_STAGES * A2_STAGE + (size_t)B2_STAGES * B_STAGE + 512 + VEC_SMEM;
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the same barrier / buffer in the pair's leader CTA (rank 0): clear the peer bit of the shared::cluster address
__device__ __forceinline__ uint32_t leader_addr(const void *p)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
    return r;
}

__device__ __forceinline__ void tma2_load_4d(void *dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void *dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_tf32_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.b32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}
// arrive (once all MMAs issued so far are complete) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__host__ __device__ constexpr uint32_t instr_desc2(int n)     // as instr_desc, M = 256 across the CTA pair
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

struct PairTile {
    long long plane;
    int y0, x0, n0;       // y0 = first of the pair's two rows
};
// pair tiles: n-tile fastest, then 128-column tile, then row pair, then plane
__device__ __forceinline__ PairTile decode_pair(const TcParams &p, long long t)
{
    PairTile r;
    r.n0 = (int)(t % p.n_tiles); t /= p.n_tiles;
    r.x0 = (int)(t % p.tiles_x) * BLOCK_M; t /= p.tiles_x;
    r.y0 = (int)(t % p.tiles_y) * 2; t /= p.tiles_y;
    r.plane = t;
    return r;
}

template <int NT, int ACTK, bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    conv_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, TcParams p,
                     const float *__restrict__ bias, const float *__restrict__ slope, const float *__restrict__ mul,
                     const float *__restrict__ residual, float *__restrict__ y)
{
    using C = Cfg2<NT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *a_base = smem;
    unsigned char *b_base = smem + (size_t)A2_STAGES * A2_STAGE;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(b_base + (size_t)B2_STAGES * C::B_STAGE);
    uint64_t *a_empty = a_full + A2_STAGES;
    uint64_t *b_full = a_empty + A2_STAGES;
    uint64_t *b_empty = b_full + B2_STAGES;
    uint64_t *acc_full = b_empty + B2_STAGES;       // [2]
    uint64_t *acc_empty = acc_full + 2;             // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    float *s_bias = reinterpret_cast<float *>(b_base + (size_t)B2_STAGES * C::B_STAGE + 512);
    float *s_slope = s_bias + MAX_CO_STAGED;
    float *s_stage = s_slope + MAX_CO_STAGED;
    stage_channel_vectors(s_bias, s_slope, bias, slope, p.Co, p.d2w);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const long long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_w);
        for (int s = 0; s < A2_STAGES; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < B2_STAGES; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; s++) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 8);           // four epilogue warps in each CTA of the pair
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int kblocks = p.Ci / BLOCK_K;

    if (warp == 0) {
        // ===================================================================================== TMA producer (both CTAs)
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (long long t = cluster_id; t < p.total_tiles; t += n_clusters) {
                PairTile tl = decode_pair(p, t);
                if (tl.x0 >= p.wl_out[(int)(tl.plane % p.npart)]) continue;
                for (int kb = 0; kb < kblocks; kb++) {
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * A2_BYTES);
                    // halo tile: input rows y .. y+2, columns x0 .. x0+129 of the (pre-padded) plane, 32 channels
                    tma2_load_4d(a_base + (size_t)sa * A2_STAGE, &map_x, leader_addr(&a_full[sa]), kb * BLOCK_K, tl.x0, tl.y0 + (int)rank,
                                 (int)tl.plane);
                    if (++sa == A2_STAGES) { sa = 0; pa ^= 1; }
                    for (int tap = 0; tap < 9; tap++) {
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        if (rank == 0) mbar_expect_tx(&b_full[sb], 2 * C::B_HALF);
                        tma2_load_3d(b_base + (size_t)sb * C::B_STAGE, &map_w, leader_addr(&b_full[sb]), kb * BLOCK_K,
                                     tl.n0 * NT + (int)rank * (NT / 2), tap);
                        if (++sb == B2_STAGES) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================================== MMA issuer (leader CTA)
        // ONE thread runs the whole loop (no per-MMA elect / reconvergence), the tap loop is fully unrolled so every
        // descriptor is `stage base + compile-time constant`: the issue loop, not L2 or the tensor pipe, was the
        // limiter of the first version (ncu source page, profiles/r1b_*).
        if (rank == 0 && lane == 0) {
            int sa = 0, sb = 0, acc = 0;
            uint32_t pa = 0, pb = 0, acc_phase = 0;
            constexpr uint32_t idesc = instr_desc2(NT);
            const uint64_t desc_hi = smem_desc(0, 16, 1024) & 0xFFFFFFFF00000000ull;
            const uint32_t a_lo0 = (smem_u32(a_base) >> 4) & 0x3fff;
            const uint32_t b_lo0 = (smem_u32(b_base) >> 4) & 0x3fff;
            long long w_acc = 0, w_a = 0, w_b = 0;
            const long long t_begin = DBG ? clock64() : 0;
            for (long long t = cluster_id; t < p.total_tiles; t += n_clusters) {
                PairTile tl = decode_pair(p, t);
                if (tl.x0 >= p.wl_out[(int)(tl.plane % p.npart)]) continue;
                long long t0 = DBG ? clock64() : 0;
                mbar_wait_cluster(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                long long t1 = DBG ? clock64() : 0;
                w_acc += t1 - t0;
                const uint32_t tmem_d = tmem_base + acc * NT;
                for (int kb = 0; kb < kblocks; kb++) {
                    if (DBG) t0 = clock64();
                    mbar_wait(&a_full[sa], pa);
                    if (DBG) t1 = clock64();
                    w_a += t1 - t0;
                    const uint32_t a_lo = a_lo0 + (uint32_t)sa * (A2_STAGE >> 4);
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        if (DBG) t0 = clock64();
                        mbar_wait(&b_full[sb], pb);
                        tc_fence_after();
                        if (DBG) t1 = clock64();
                        w_b += t1 - t0;
                        const uint32_t b_lo = b_lo0 + (uint32_t)sb * (C::B_STAGE >> 4);
                        const int r0 = (tap / 3) * HALO_W + (tap % 3);          // compile-time after unrolling
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / UMMA_K; kk++) {
                            const uint64_t ad = desc_hi | (uint64_t)(a_lo + (uint32_t)((r0 * BLOCK_K * 4 + kk * UMMA_K * 4) >> 4));
                            const uint64_t bd = desc_hi | (uint64_t)(b_lo + (uint32_t)((kk * UMMA_K * 4) >> 4));
                            if (tap == 0 && kk == 0) umma2_tf32(tmem_d, ad, bd, idesc, kb != 0);
                            else umma2_tf32_acc(tmem_d, ad, bd, idesc);
                        }
                        umma2_commit(&b_empty[sb]);
                        if (++sb == B2_STAGES) { sb = 0; pb ^= 1; }
                    }
                    umma2_commit(&a_empty[sa]);
                    if (++sa == A2_STAGES) { sa = 0; pa ^= 1; }
                }
                umma2_commit(&acc_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (DBG && p.dbg) {
                long long *d = p.dbg + 4 * cluster_id;
                d[0] = clock64() - t_begin; d[1] = w_acc; d[2] = w_a; d[3] = w_b;
            }
        }
    } else {
        // ===================================================================================== epilogue warps (both CTAs)
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long ep_wait = 0, ep_work = 0, ep_ld = 0;
        for (long long t = cluster_id; t < p.total_tiles; t += n_clusters) {
            PairTile tl = decode_pair(p, t);
            const int wl = p.wl_out[(int)(tl.plane % p.npart)];
            const int oy = tl.y0 + (int)rank;
            const int ox0 = tl.x0 + q * 32;
            const bool live = tl.x0 < wl;
            long long e0 = 0, e1 = 0;
            if (DBG) e0 = clock64();
            if (live) {
                mbar_wait(&acc_full[acc], acc_phase);
                tc_fence_after();
            }
            if (DBG) e1 = clock64();
            epilogue_warp<NT, ACTK, DBG>(p, tl.plane, oy, ox0, tl.n0, live, wl, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT),
                                   s_stage + q * 32 * EPI_PITCH, bias ? s_bias : nullptr, s_slope, mul, residual, y, &ep_ld);
            if (DBG) { ep_wait += e1 - e0; ep_work += clock64() - e1; }
            if (live) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_addr(&acc_empty[acc]));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        if (DBG && p.dbg && warp == 2 && lane == 0 && rank == 0) {
            long long *d = p.dbg + 4 * 74 + 2 * cluster_id;
            d[0] = ep_wait; d[1] = ep_work; p.dbg[6 * 74 + cluster_id] = ep_ld;
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, C::TMEM_COLS);
    }
}

// OIHW fp32 -> [tap][Co_pad][Ci], values rounded to the nearest TF32 (the MMA would otherwise truncate them)
__global__ void pack_weights_kernel(const float *__restrict__ w, float *__restrict__ out, int Co, int Ci, int kk, int co_pad, int d2w)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)kk * co_pad * Ci;
    if (idx >= total) return;
    int ci = (int)(idx % Ci);
    int co = (int)((idx / Ci) % co_pad);
    int tap = (int)(idx / Ci / co_pad);
    float v = 0.f;
    if (co < Co) {
        const int cs = d2w ? 4 * (co % (Co >> 2)) + co / (Co >> 2) : co;      // depth-to-space column order (see TcParams::d2w)
        v = w[((long long)cs * Ci + ci) * kk + tap];
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
        v = __uint_as_float(r);
    }
    out[idx] = v;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

int n_tile_for(int Co) { return Co <= 16 ? 16 : (Co % 192 == 0 ? 192 : (Co % 96 == 0 ? 96 : 0)); }

struct PackedWeights {
    const float *src = nullptr;
    float *dst = nullptr;
    int Co = 0, Ci = 0, k = 0, co_pad = 0;
    unsigned long long stamp = 0;
};

}  // namespace

// Weight repacking is cached per (pointer, shape, content stamp); the stamp is the caller's version counter.
static std::mutex g_pack_mutex;
static PackedWeights g_pack_cache[512];
static int g_pack_next = 0;

static long long pack_weights(const float *d_w, float *d_out, int Co, int Ci, int k, int d2w, void *stream);

extern "C" long long pcx_conv_pack_weights(const float *d_w, float *d_out, int Co, int Ci, int k, void *stream)
{
    return pack_weights(d_w, d_out, Co, Ci, k, 0, stream);
}

extern "C" long long pcx_conv_pack_weights_d2w(const float *d_w, float *d_out, int Co, int Ci, int k, void *stream)
{
    if (Co % 4 != 0 || n_tile_for(Co) != Co / 4) {
        pcx_set_error("fused depth-to-space needs Co / 4 to be one N tile (96 or 192 channels), got Co=%d", Co);
        return PCX_EINVAL;
    }
    return pack_weights(d_w, d_out, Co, Ci, k, 1, stream);
}

static long long pack_weights(const float *d_w, float *d_out, int Co, int Ci, int k, int d2w, void *stream)
{
    int nt = n_tile_for(Co);
    if (nt == 0 || Ci % BLOCK_K != 0 || (k != 1 && k != 3)) {
        pcx_set_error("tensor-core conv supports Co <= 16 or a multiple of 96, Ci a multiple of 32, k in {1,3} (got Co=%d Ci=%d k=%d)", Co, Ci, k);
        return PCX_EINVAL;
    }
    int co_pad = (Co + nt - 1) / nt * nt;
    long long total = (long long)k * k * co_pad * Ci;
    if (d_out == nullptr) return total;
    if (d_w == nullptr) { pcx_set_error("null weights"); return PCX_EINVAL; }
    pack_weights_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(d_w, d_out, Co, Ci, k * k, co_pad, d2w);
    PCX_LAUNCHED();
    return total;
}

template <int NT, int ACTK, bool SQ = false>
static int launch_tc_v(const CUtensorMap &mx, const CUtensorMap &mw, const TcParams &p, const float *bias, const float *slope,
                     const float *mul, const float *residual, float *y, cudaStream_t s)
{
    using C = Cfg<NT>;
    static PcxDeviceOnce once;
    PCX_ONCE_PER_DEVICE(once) PCX_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT, ACTK, SQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    long long grid = p.total_tiles < pcx_sm_count() ? p.total_tiles : pcx_sm_count();
    conv_tc_kernel<NT, ACTK, SQ><<<(unsigned)grid, TC_THREADS + (SQ ? SQ_WARPS * 32 : 0), C::SMEM, s>>>(mx, mw, p, bias, slope, mul, residual, y);
    PCX_LAUNCHED();
    return PCX_OK;
}


template <int NT, int ACTK, bool DBG>
static int launch_pair_v(const CUtensorMap &mx, const CUtensorMap &mw, const TcParams &p, const float *bias, const float *slope,
                       const float *mul, const float *residual, float *y, cudaStream_t s)
{
    using C = Cfg2<NT>;
    static PcxDeviceOnce once;
    PCX_ONCE_PER_DEVICE(once) {
        PCX_CUDA(cudaFuncSetAttribute(conv_pair_kernel<NT, ACTK, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pcx_sm_count() / 2 * 2);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = C::SMEM;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, conv_pair_kernel<NT, ACTK, DBG>, &cfg);
        if (e != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = pcx_sm_count() / 2; }
        once.slot[pcx_current_device()] = n;
    }
    const int max_clusters = once.slot[pcx_current_device()];
    long long clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
    conv_pair_kernel<NT, ACTK, DBG><<<(unsigned)(2 * clusters), NUM_THREADS, C::SMEM, s>>>(mx, mw, p, bias, slope, mul, residual, y);
    PCX_LAUNCHED();
    return PCX_OK;
}


template <int NT>
static int launch_tc(const CUtensorMap &mx, const CUtensorMap &mw, const TcParams &p, const float *bias, const float *slope,
                     const float *mul, const float *residual, float *y, cudaStream_t s)
{
    if (p.square) {         // GDN / IGDN layers only (pcx_conv2d_tc checks)
        return p.act == 3 ? launch_tc_v<NT, 3, true>(mx, mw, p, bias, slope, mul, residual, y, s)
                          : launch_tc_v<NT, 4, true>(mx, mw, p, bias, slope, mul, residual, y, s);
    }
    switch (p.act) {
    case 2: return launch_tc_v<NT, 2>(mx, mw, p, bias, slope, mul, residual, y, s);
    case 3: return launch_tc_v<NT, 3>(mx, mw, p, bias, slope, mul, residual, y, s);
    case 4: return launch_tc_v<NT, 4>(mx, mw, p, bias, slope, mul, residual, y, s);
    default:    // ACTK 5 = plain / PReLU epilogue that also carries a gate operand (not used by the codec's layers)
        return mul ? launch_tc_v<NT, 5>(mx, mw, p, bias, slope, mul, residual, y, s) : launch_tc_v<NT, 0>(mx, mw, p, bias, slope, mul, residual, y, s);
    }
}
template <int NT, bool DBG = false>
static int launch_pair(const CUtensorMap &mx, const CUtensorMap &mw, const TcParams &p, const float *bias, const float *slope,
                       const float *mul, const float *residual, float *y, cudaStream_t s)
{
    // transcendental epilogues only occur on 1x1 layers (sigmoid gates, GDN) - the pair kernel is 3x3 only
    return launch_pair_v<NT, 0, DBG>(mx, mw, p, bias, slope, mul, residual, y, s);
}

static int pair_mode()
{
    // PCX_TC_PAIR: 0 = never use the CTA-pair kernel, 1 (default) = 3x3 stride-1 layers with Wo >= 64
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("PCX_TC_PAIR");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

int pcx_conv2d_tc(const pcx_conv_desc *desc, const float *d_x, const float *d_w, const float *d_bias, const float *d_slope,
                  const float *d_mul, const float *d_residual, float *d_y, void *stream)
{
    const pcx_conv_desc &d = *desc;
    cudaStream_t s = (cudaStream_t)stream;
    const int nt = n_tile_for(d.Co);
    PCX_REQUIRE(nt != 0 && d.Ci % BLOCK_K == 0, "tensor-core conv needs Co <= 16 or a multiple of 96 and Ci a multiple of 32 (Co=%d Ci=%d); use impl=1", d.Co, d.Ci);
    PCX_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_y) & 15) == 0, "tensor-core conv needs 16-byte aligned NHWC buffers");
    PCX_REQUIRE(d.Co % 4 == 0 && d.Co <= MAX_CO_STAGED, "tensor-core conv needs Co %% 4 == 0 and Co <= %d (got %d)", MAX_CO_STAGED, d.Co);
    const int bw = d.Wo >= 512 ? 128 : (d.Wo >= 256 ? 64 : 32);
    const int bh = BLOCK_M / bw;
    EncodeTiledFn enc = encode_tiled();
    PCX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");

    // ---- weights: impl 2 = d_w is already in the packed [tap][Co_pad][Ci] layout (pcx_conv_pack_weights);
    //      impl 0 = d_w is the caller's OIHW tensor, repacked into a cached scratch buffer on every call (the
    //      tensor may have been updated in place; the repack is <= 5 MB)
    const int co_pad = (d.Co + nt - 1) / nt * nt;
    float *packed = nullptr;
    PCX_REQUIRE(d.impl != 3 || (d.Co % 4 == 0 && nt == d.Co / 4), "fused depth-to-space needs Co / 4 to be one N tile (96 or 192 channels), got Co=%d", d.Co);
    if (d.impl == 2 || d.impl == 3) {
        PCX_REQUIRE((reinterpret_cast<uintptr_t>(d_w) & 15) == 0, "packed weights must be 16-byte aligned");
        packed = const_cast<float *>(d_w);
    } else {
        {
            std::lock_guard<std::mutex> lock(g_pack_mutex);
            for (auto &e : g_pack_cache)
                if (e.src == d_w && e.Co == d.Co && e.Ci == d.Ci && e.k == d.k) { packed = e.dst; break; }
            if (!packed) {
                PackedWeights &e = g_pack_cache[g_pack_next];
                g_pack_next = (g_pack_next + 1) % 512;
                if (e.dst) cudaFree(e.dst);
                size_t n = (size_t)d.k * d.k * co_pad * d.Ci;
                PCX_CUDA(cudaMalloc(&e.dst, n * sizeof(float)));
                e.src = d_w; e.Co = d.Co; e.Ci = d.Ci; e.k = d.k; e.co_pad = co_pad;
                packed = e.dst;
            }
        }
        long long rc = pcx_conv_pack_weights(d_w, packed, d.Co, d.Ci, d.k, stream);
        if (rc < 0) return (int)rc;
    }

    PCX_REQUIRE(d.square_input == 0 || (d.k == 1 && (d.act == 3 || d.act == 4)), "square_input is the GDN / IGDN form: 1x1, act 3 or 4 (k=%d act=%d)", d.k, d.act);
    const bool pair = pair_mode() != 0 && d.k == 3 && d.stride == 1 && d.Wo >= 64 && d.act <= 1 && d_mul == nullptr;
    const cuuint64_t plane_rows = (cuuint64_t)(d.in_plane_rows > 0 ? d.in_plane_rows : d.Hi);

    // ---- tensor maps
    CUtensorMap mx, mw;
    if (pair) {
        // halo tile of the CTA-pair kernel: 32 channels x 130 columns x 3 rows
        const long long planes = (long long)d.N * d.npart;
        cuuint64_t dims[4] = {(cuuint64_t)d.Ci, (cuuint64_t)d.in_pitch, (cuuint64_t)d.Hi, (cuuint64_t)planes};
        cuuint64_t strides[3] = {(cuuint64_t)d.Ci * 4, (cuuint64_t)d.Ci * d.in_pitch * 4, (cuuint64_t)d.Ci * d.in_pitch * plane_rows * 4};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)HALO_W, 3, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(d_x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PCX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo tile) failed with %d", (int)r);
    } else {
        const long long planes = (long long)d.N * d.npart;
        cuuint64_t dims[4] = {(cuuint64_t)d.Ci, (cuuint64_t)d.in_pitch, (cuuint64_t)d.Hi, (cuuint64_t)planes};
        cuuint64_t strides[3] = {(cuuint64_t)d.Ci * 4, (cuuint64_t)d.Ci * d.in_pitch * 4, (cuuint64_t)d.Ci * d.in_pitch * plane_rows * 4};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)(bw * d.stride), (cuuint32_t)(bh * d.stride), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(d_x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PCX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activations) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)d.Ci, (cuuint64_t)co_pad, (cuuint64_t)(d.k * d.k)};
        cuuint64_t strides[2] = {(cuuint64_t)d.Ci * 4, (cuuint64_t)d.Ci * co_pad * 4};
        cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)(pair ? nt / 2 : nt), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, packed, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PCX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    }

    TcParams p;
    p.planes = d.N * d.npart; p.npart = d.npart;
    p.Ci = d.Ci; p.Co = d.Co; p.Ho = d.Ho; p.Wo = d.Wo;
    p.out_rows = d.out_rows; p.out_pitch = d.out_pitch; p.out_y0 = d.out_y0; p.out_x0 = d.out_x0;
    p.aux_rows = d.aux_rows; p.aux_pitch = d.aux_pitch; p.aux_y0 = d.aux_y0; p.aux_x0 = d.aux_x0;
    p.k = d.k; p.stride = d.stride; p.act = d.act;
    p.square = d.square_input;
    p.d2w = d.impl == 3 ? 1 : 0;
    p.bw = bw; p.bh = bh;
    p.tiles_x = (d.Wo + bw - 1) / bw;
    p.tiles_y = (d.Ho + bh - 1) / bh;
    p.n_tiles = co_pad / nt;
    p.co_pad = co_pad;
    p.total_tiles = (long long)p.planes * p.tiles_y * p.tiles_x * p.n_tiles;
    for (int i = 0; i < PCX_MAX_PART; i++) p.wl_out[i] = d.wl_out[i];

    p.dbg = nullptr;
    if (pair) {
        static long long *dbg = nullptr;
        if (getenv("PCX_TC_DEBUG")) {
            if (!dbg) { cudaMalloc(&dbg, 8 * 128 * sizeof(long long)); }
            p.dbg = dbg;
        }
        p.bw = BLOCK_M; p.bh = 1;
        p.tiles_x = (d.Wo + BLOCK_M - 1) / BLOCK_M;
        p.tiles_y = (d.Ho + 1) / 2;
        p.total_tiles = (long long)p.planes * p.tiles_y * p.tiles_x * p.n_tiles;
        if (p.dbg) {
            int rc = nt == 192 ? launch_pair<192, true>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s)
                               : (nt == 96 ? launch_pair<96, true>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s)
                                           : launch_pair<16, true>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s));
            static int printed = 0;
            if (printed++ < 3) {
                long long h[7 * 74];
                cudaStreamSynchronize(s);
                cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
                for (int c = 0; c < 74; c += 18)
                    fprintf(stderr, "[pcx dbg] cluster %2d: total %lld clk, wait acc_empty %lld, a_full %lld, b_full %lld | epilogue warp: wait acc_full %lld, work %lld of which tcgen05.ld %lld\n",
                            c, h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3], h[4 * 74 + 2 * c], h[4 * 74 + 2 * c + 1], h[6 * 74 + c]);
            }
            return rc;
        }
        if (nt == 192) return launch_pair<192>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s);
        if (nt == 96) return launch_pair<96>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s);
        return launch_pair<16>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s);
    }
    if (nt == 192) return launch_tc<192>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s);
    if (nt == 96) return launch_tc<96>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s);
    return launch_tc<16>(mx, mw, p, d_bias, d_slope, d_mul, d_residual, d_y, s);
}
