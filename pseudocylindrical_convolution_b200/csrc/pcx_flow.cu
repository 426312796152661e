// pcx_flow.cu - the decoder's wavefront loop (EntDecoder.forward, pseudo_codec.py:145-160) as ONE persistent dataflow kernel.
//
// The reference runs ~32 launches and two blocking copies per wavefront step; pcx_ctx.cu's fused step kernel made that one
// cooperative launch per step with 13 grid barriers (~2.5 us each) and a launch + stream synchronisation + host decode round
// trip (~26 us) - 131 us per step at 512x1024, all latency.  Here the whole decode is one launch and nothing in it is a barrier:
//
//   * every scratch scalar of the 12 masked layers is written exactly once per decode, into buffers pre-filled with a NaN
//     payload that no arithmetic produces (FLOW_SENTINEL): a value is its own "ready" flag.  A warp that needs a scalar of the
//     current step polls that one word in L2 until its producer has stored it (step_conv_phase<GI, true>); layer L of a cell
//     starts the moment its 5x5 neighbours of layer L-1 exist - no grid-wide rendezvous, no launch boundary.  Arithmetic (FFMA
//     chains in ascending channel-group order, the reference's fold tree) is the shared step_conv_phase code: bit-identical CDFs.
//   * images are independent pipelines: block b works for image b % nimg only, so the images of a batch drift apart and one
//     image's host round trip is hidden behind the others' device work (no lock-step over the batch).
//   * device <-> host through mapped pinned memory only, no CUDA call inside the loop: a CDF row leaves as ONE 16-byte store
//     (cum[1..7] as uint16 + a 16-bit step tag; cum[0] = 0 and cum[8] = 65536 are constants) and the host decodes rows as their
//     tags appear; a decoded symbol returns as one 32-bit word (step tag << 8 | symbol) that the image's leader block polls.
//
// Deadlock freedom: all blocks are co-resident (cooperative launch); a block executes its (layer, run) items in layer order
// and an item of layer L waits only for scalars of layer L-1 (or for the host, which waits only for layer 11 of the previous
// step), so the unfinished item with the smallest (step, layer) can always proceed.  Every wait checks an abort word and a
// time-out, so a corrupt bitstream or a lost host cannot hang the device.
#include "pcx_ctx_step.cuh"
#include "pcx_flow.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string.h>
#include <thread>
#include <vector>
#include <sched.h>
#include <immintrin.h>

namespace {

struct FlowNet {
    StepNet net;                  // net.prev / net.cdf / net.bar are unused here
    int nsteps, rows_cap, blocks_per_img;
    unsigned tag_salt;            // per-call salt of the row / symbol tags: a word left over from an earlier decode never matches
    const int *start;             // device copy of the plane prefix of the wavefront order (Hf + W entries)
    const int4 *sched;            // per step: first plane, planes, run length, runs (per image)
    uint4 *rows;                  // mapped host memory: (nimg, rows_cap) CDF rows, 7 x uint16 boundaries + uint16 tag (step + 1)
    const unsigned *symw;         // mapped host memory: (nimg, rows_cap) decoded symbols, (step + 1) << 8 | symbol
    unsigned *host_ctl;           // mapped host memory: [0] host -> device abort request, [1] device -> host error word
    unsigned *dev_ctl;            // device memory: [0] abort flag, [16 + img] = steps whose input symbols are in place
    unsigned long long timeout_ns;
    const float *packed;          // weight images (flow_pack_weights_kernel), wbuf floats per (layer, net, channel group)
    unsigned long long *trace;    // optional (PCX_WAVE_TRACE): per step 4 globaltimer stamps of image 0's leader block
};

__global__ void fill_u32_kernel(unsigned *p, size_t n, unsigned v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// Weight rows in the layout the step kernels keep in shared memory, once per decode: packed[((L * nb + b) * G + tc)] is the
// image of outputs tc*3 .. tc*3+2 of net b in layer L - [ci * 25 + tap][og] floats (4 per entry, the three outputs + padding),
// then bias and slope at 4 * wstride.  Staging a (layer, net, channel group) for a run is then a straight 16-byte cp.async
// copy; the gather form (step_stage_weights: 4-byte copies, two integer divisions each) cost every thread ~400 instructions
// per run, as much as the run's arithmetic when a block's runs are 8 cells long (batches).
__global__ void flow_pack_weights_kernel(const StepNet d, float *__restrict__ packed, int wbuf)
{
    const int G = d.G, Co = G * 3;
    const i64 per_layer = (i64)d.nb * G * wbuf;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < per_layer * d.nlayers; i += (i64)gridDim.x * blockDim.x) {
        const int L = (int)(i / per_layer);
        const i64 r0 = i - L * per_layer;
        const int bt = (int)(r0 / wbuf), e = (int)(r0 - (i64)bt * wbuf);
        const int b = bt / G, tc = bt - b * G;
        const StepLayer &l = d.L[L];
        const int Ci = G * l.gi, nw = Ci * 25, tail = 4 * G * 3 * 25;
        float v = 0.f;
        if (e < 4 * nw) {
            const int r = e >> 2, og = e & 3;
            if (og < 3) v = l.weight[(((i64)b * Co + tc * 3 + og) * Ci) * 25 + r];
        } else if (e >= tail && e < tail + 3) {
            v = l.bias[b * Co + tc * 3 + (e - tail)];
        } else if (e >= tail + 3 && e < tail + 6 && l.act != nullptr) {
            v = l.act[b * Co + tc * 3 + (e - tail - 3)];
        }
        packed[i] = v;
    }
}

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// the first gmax channel groups of the packed image (+ bias / slope) into a shared-memory weight buffer
__device__ __forceinline__ void flow_stage_packed(const StepNet &d, const float *packed, int wbuf, int L, int b, int tc, float *smem)
{
    const StepLayer &l = d.L[L];
    int gmax = tc + 4 + (l.constrain == 6 ? 1 : 0);
    gmax = gmax > d.G ? d.G : gmax;
    const int nvec = gmax * l.gi * 25;                                    // 16-byte entries
    const float *src = packed + (((i64)L * d.nb + b) * d.G + tc) * wbuf;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) cp_async16(smem + 4 * i, src + 4 * i);
    const int tail = 4 * d.G * 3 * 25;
    if (threadIdx.x < 2) cp_async16(smem + tail + 4 * threadIdx.x, src + tail + 4 * threadIdx.x);
}

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long flow_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// symbol word t of the leader's image for `step`: poll host memory until its tag matches
__device__ __noinline__ unsigned flow_poll_symbol(const FlowNet &f, const unsigned *p, unsigned tag)
{
    unsigned w, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        w = ld_volatile_u32(p);
        if ((w >> 8) == tag) break;
        if ((++spins & 15u) == 0) {
            if (ld_relaxed_u32(f.dev_ctl) != 0) break;
            if (ld_volatile_u32(f.host_ctl) != 0) { atomicExch(f.dev_ctl, 1u); break; }
            const unsigned long long now = flow_now();
            if (t0 == 0) t0 = now;
            else if (now - t0 > f.timeout_ns) { atomicExch(f.dev_ctl, 2u); break; }
        }
    }
    return w;
}

// Measured alternatives to the warp-per-cell task that are NOT in this file any more (bit-identical, all slower on the B200;
// numbers in profiles/r2c_flow_task_probe.txt): four lanes per task, C = 2 / 3 cells per warp sharing the weight loads, a
// member-major layout with the loads of member m + 1 issued under the FFMAs of member m, L1 prefetch of the next task's operands.
// Also measured and dropped in round 2 (DESIGN.md section 3.3 has the numbers): operands staged in shared memory with cp.async, a
// branch-free FFMA batch, three and five blocks per SM at 80 / 96 registers, and a lane-per-cell form (a lane runs the 75 chains
// of its cell, fold tree on an in-lane stack: 3.5x fewer instructions, 12x slower - one exposed L2 round trip per chain).
// NT threads per block, BPS blocks per SM, the first CACHE cells of a block's step resolved in shared memory.  Two forms are
// built: 256 x 2 (cache 64) when a layer of a step is about one task per warp (one small image: fewest blocks to rendezvous
// per weight change), 128 x 4 (cache 32) when every warp has several tasks per layer (batches, large images): smaller blocks
// couple fewer warps at the barrier that follows a weight change (entropy decode, ms, 256 x 2 -> 128 x 4: 8 x 512x1024 60.2 -> 55.9,
// 16 x 512x1024 122.5 -> 111.5, 1 x 1024x2048 47.0 -> 46.3, 1 x 2048x4096 149.7 -> 146.9; 1 x 512x1024 18.3 -> 18.4).
template <int NT, int BPS, int CACHE>
__global__ void __launch_bounds__(NT, BPS) wave_flow_kernel(const __grid_constant__ FlowNet f)
{
    extern __shared__ __align__(16) float step_ws[];  // 2 buffers of 4 * wstride weights + 8 (bias, slope)
    const StepNet &d = f.net;
    const int img = blockIdx.x % d.nimg, bi = blockIdx.x / d.nimg, B = f.blocks_per_img;
    const int tid = threadIdx.x;
    const int h = d.h, W = d.W, pad = d.pad, G = d.G;
    const int wstride = G * 3 * 25;
    const int wbuf = 4 * wstride + 8;                 // floats per weight buffer / packed image
    const int *__restrict__ start = f.start;
    FlowCtl ctl;
    ctl.abort_flag = f.dev_ctl;
    ctl.timeout_ns = f.timeout_ns;
    unsigned *ready = f.dev_ctl + 16 + img;
    __shared__ StepChunk s_chunk[STEP_MAX_RUNS];
    __shared__ int s_cache_base[STEP_MAX_RUNS + 1];
    __shared__ StepCellRec s_cell[CACHE];
    __shared__ StepTapOff s_tap[CACHE * 25];
    __shared__ unsigned s_abort;
    const bool tracer = f.trace != nullptr && blockIdx.x == 0 && tid == 0;

    for (int step = 0; step <= f.nsteps; step++) {
        int pfirst = 0, pcount = 0;                    // window of step - 1: its symbols enter the network input now
        if (step > 0) {
            const int4 ps = f.sched[step - 1];
            pfirst = start[ps.x];
            pcount = start[ps.x + ps.y] - pfirst;
        }
        int p0 = 0, np = 0, S = 1, nchunk = 0, first = 0, count = 0;
        if (step < f.nsteps) {
            const int4 sc = f.sched[step];
            p0 = sc.x; np = sc.y; S = sc.z; nchunk = sc.w;
            first = start[p0];
            count = start[p0 + np] - first;
        }
        // the block's runs: a CONTIGUOUS range of the step's run list (ordered plane, net, run) - consecutive runs mostly share
        // their (net, plane), i.e. their weight rows, which are then staged once
        const int per = (nchunk + B - 1) / B;
        const int c_first = bi * per;
        const int nmy = c_first < nchunk ? (nchunk - c_first < per ? nchunk - c_first : per) : 0;
        const int nitems = nmy * d.nlayers;
        if (tracer) f.trace[(size_t)step * 16 + 0] = flow_now();
        if (tid < nmy && tid < STEP_MAX_RUNS) s_chunk[tid] = step_chunk_of(start, c_first + tid, p0, np, S, d.nb, 1);
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int i = 0; i < nmy && i < STEP_MAX_RUNS; i++) { s_cache_base[i] = acc; acc += s_chunk[i].ncell; }
            s_cache_base[nmy < STEP_MAX_RUNS ? nmy : STEP_MAX_RUNS] = acc;
        }
        __syncthreads();
        {
            // (cell, tap) geometry of the block's first CACHE cells, once for all layers: warp per cell, lane per tap
            const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
            const int nruns = nmy < STEP_MAX_RUNS ? nmy : STEP_MAX_RUNS;
            const int ncached = s_cache_base[nruns] < CACHE ? s_cache_base[nruns] : CACHE;
            for (int c = warp; c < ncached; c += nwarp) {
                int ci = 0;
                while (ci + 1 < nruns && s_cache_base[ci + 1] <= c) ci++;
                const StepChunk ch = s_chunk[ci];
                const int k = c - s_cache_base[ci];
                const int pn = ch.net * d.nimg + img;
                const int4 cinfo = d.cell[start[ch.plane] + ch.rem0 + k];
                if (lane == 0) s_cell[c] = StepCellRec{pn, cinfo.z, cinfo.w, cinfo.x};
                if (lane < 25) {
                    const int kw = lane % 5, kh = lane / 5;
                    s_tap[c * 25 + lane] = step_resolve_off(d, pn, cinfo.z, cinfo.w + kh - 2, cinfo.x + kw - 2);
                }
            }
        }
        if (nitems > 0) {
            const StepChunk c0 = s_chunk[0];
            flow_stage_packed(d, f.packed, wbuf, 0, c0.net, step - c0.plane, step_ws);
        }
        cp_async_commit();

        // ---- DInput2: the image's leader block takes the symbols of step - 1 from the host and publishes the step
        if (bi == 0) {
            if (pcount > 0) {
                const i64 ih = h + 2 * pad, iw = W + 2 * pad;
                const int cp0 = d.L[0].cp_in;
                const i64 rep_stride = (i64)d.nimg * d.npart * ih * iw * cp0;
                float *out = const_cast<float *>(d.L[0].in);
                const unsigned *words = f.symw + (size_t)img * f.rows_cap;
                const unsigned tag = 0x800000u | ((f.tag_salt & 0x7fu) << 16) | ((unsigned)step & 0xffffu);      // step = (decoded step) + 1
                for (int base = 0; base < pcount; base += 4 * NT) {
                    unsigned w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {                  // four PCIe reads in flight per thread
                        const int t = base + j * NT + tid;
                        w[j] = t < pcount ? ld_volatile_u32(words + t) : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int t = base + j * NT + tid;
                        if (t >= pcount) continue;
                        if ((w[j] >> 8) != tag) w[j] = flow_poll_symbol(f, words + t, tag);
                        const int4 ci = d.cell[pfirst + t];
                        const int tw = ci.x, hp = ci.y, g = ci.z, th = ci.w;
                        const int tc = step - 1 - tw - hp;
                        const float v = __fadd_rn((float)(w[j] & 0xffu), d.input_bias);
                        const i64 i = ((((i64)img * d.npart + g) * ih + th + pad) * iw + tw + pad) * cp0 + tc;
                        for (int r = 0; r < d.nb; r++) __stcg(out + i + r * rep_stride, v);
                        d.sym_nchw[((((i64)img * d.npart + g) * G + tc) * ih + th + pad) * iw + tw + pad] = v;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ready), "r"((unsigned)step + 1u) : "memory");
                s_abort = ld_relaxed_u32(f.dev_ctl);
            }
        } else if (tid == 0) {
            unsigned spins = 0, a = 0;
            unsigned long long t0 = 0;
            while ((int)(ld_relaxed_u32(ready) - ((unsigned)step + 1u)) < 0) {
                if ((++spins & 63u) == 0) {
                    a = ld_relaxed_u32(f.dev_ctl);
                    if (a != 0) break;
                    const unsigned long long now = flow_now();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > f.timeout_ns) { atomicExch(f.dev_ctl, 2u); a = 2; break; }
                }
            }
            // one acquire fence after the wait (an acquire load inside the loop would invalidate this SM's L1 at every poll):
            // scalars of earlier steps are read through L1 below, this drops the lines cached while they still held the sentinel
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            s_abort = a != 0 ? a : ld_relaxed_u32(f.dev_ctl);
        }
        __syncthreads();
        if (s_abort != 0) {
            if (blockIdx.x == 0 && tid == 0) {
                f.host_ctl[1] = s_abort;
                __threadfence_system();
            }
            return;
        }
        if (tracer) f.trace[(size_t)step * 16 + 1] = flow_now();

        // ---- the masked layers: no barrier between them, consumers poll the scalars they need (step_conv_phase<GI, true>)
        // Weight buffers: an item whose (layer, net, plane) equals the previous item's reuses the rows in place - no copy and NO
        // block barrier, the warps of the block drift apart over such items.  Only an item with new rows costs one barrier: it
        // says that every thread's copies have landed and - every warp having finished the previous item - that the OTHER buffer
        // is free, so the rows of the next different item are staged right after it, under this item's arithmetic.
        int it = 0, buf = 0;
        bool fresh = true;                                 // the current buffer was filled since the last barrier
        for (int L = 0; L < d.nlayers; L++) {
            for (int ci = 0; ci < nmy; ci++, it++) {
                const StepChunk ch = ci < STEP_MAX_RUNS ? s_chunk[ci] : step_chunk_of(start, c_first + ci, p0, np, S, d.nb, 1);
                if (fresh) {
                    cp_async_wait<0>();
                    __syncthreads();
                }
                bool restage = false;
                if (it + 1 < nitems) {
                    const int nci = ci + 1 < nmy ? ci + 1 : 0, nL = ci + 1 < nmy ? L : L + 1;
                    const StepChunk nc = nci < STEP_MAX_RUNS ? s_chunk[nci] : step_chunk_of(start, c_first + nci, p0, np, S, d.nb, 1);
                    restage = nL != L || nc.net != ch.net || nc.plane != ch.plane;
                    if (restage) {
                        flow_stage_packed(d, f.packed, wbuf, nL, nc.net, step - nc.plane, step_ws + (buf ^ 1) * wbuf);
                        cp_async_commit();
                    }
                }
                const float *ws = step_ws + buf * wbuf;
                const int cbase = ci < STEP_MAX_RUNS ? s_cache_base[ci] : CACHE;
                if (d.L[L].gi == 1) step_conv_phase<1, true>(d, d.L[L], step, ch, start, ws, wstride, s_tap, s_cell, cbase, img, &ctl, CACHE);
                else step_conv_phase<3, true>(d, d.L[L], step, ch, start, ws, wstride, s_tap, s_cell, cbase, img, &ctl, CACHE);
                fresh = restage;
                if (restage) buf ^= 1;
            }
            if (tracer && L < 12) f.trace[(size_t)step * 16 + 2 + L] = flow_now();
        }
        cp_async_wait<0>();
        if (tracer) f.trace[(size_t)step * 16 + 14] = flow_now();

        // ---- DExtract2Batch + GMM table: row k of the image's window, one 16-byte store to the host per row
        if (count > 0) {
            const StepLayer &last = d.L[d.nlayers - 1];
            const i64 net_stride = (i64)d.nimg * d.npart * h * W * last.cp_out;
            uint4 *rows = f.rows + (size_t)img * f.rows_cap;
            // eight lanes per row: lane j builds boundary j + 1 (the expensive part: per mixture component a division, an erf and
            // two double FMAs), the seven values meet through shuffles and every lane runs the cheap fix-up; lane 0 stores.  The
            // row is on the serial path of the step - the host cannot start before it - so its latency, not its work, counts.
            const int sub = tid & 7, grp = tid >> 3;
            const int rounds = (count - bi + B * (NT >> 3) - 1) / (B * (NT >> 3));        // the same for every thread of the block
            for (int it = 0; it < rounds; it++) {
                const int k = bi + B * (grp + it * (NT >> 3));
                const bool valid = k < count;
                float cpt = 0.f;
                if (valid) {
                    const int4 ci = d.cell[first + k];
                    const int tw = ci.x, hp = ci.y, g = ci.z, th = ci.w;
                    const int tc = step - tw - hp;
                    const float *pp = last.out + ((((i64)img * d.npart + g) * h + th) * W + tw) * last.cp_out + tc * 3;
                    float v[9];
#pragma unroll
                    for (int j = 0; j < 9; j++) v[j] = flow_ld(pp + (j / 3) * net_stride + (j % 3));   // nets: logits, delta, mean
#pragma unroll
                    for (int j = 0; j < 9; j++)
                        if (__float_as_uint(v[j]) == FLOW_SENTINEL) v[j] = flow_poll(pp + (j / 3) * net_stride + (j % 3), &ctl);
                    float w[3] = {v[0], v[1], v[2]}, dl[3] = {v[3], v[4], v[5]}, mu[3] = {v[6], v[7], v[8]};
                    gmm_prepare<3>(w, dl, 3, d.gmm_beta);
                    cpt = gmm_boundary<3>(w, dl, mu, 3, sub < 7 ? sub + 1 : 7, d.gmm_bias, d.gmm_total, 0);
                }
                float c[9];
#pragma unroll
                for (int j = 1; j < 8; j++) c[j] = __shfl_sync(0xffffffffu, cpt, (tid & 24) + j - 1);
                gmm_fixup<8>(c, 8, d.gmm_total);
                if (valid && sub == 0) {
                    uint4 r;
                    r.x = (unsigned)(int)c[1] | ((unsigned)(int)c[2] << 16);
                    r.y = (unsigned)(int)c[3] | ((unsigned)(int)c[4] << 16);
                    r.z = (unsigned)(int)c[5] | ((unsigned)(int)c[6] << 16);
                    r.w = (unsigned)(int)c[7] | ((0x8000u | ((f.tag_salt + (unsigned)step) & 0x7fffu)) << 16);
                    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(rows + k), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w)
                                 : "memory");
                }
            }
        }
        if (tracer) f.trace[(size_t)step * 16 + 15] = flow_now();
        __syncthreads();                               // the step's shared tables are rebuilt next
    }
}

// ---------------------------------------------------------------------------------------------- host side
struct FlowState {                  // per device, created on first use, guarded by `mu` (one decode per device at a time)
    std::mutex mu;
    float *scratch = nullptr;
    size_t scratch_bytes = 0;
    unsigned *dev_ctl = nullptr;
    int *d_start = nullptr;
    int4 *d_sched = nullptr;
    size_t start_cap = 0, sched_cap = 0;
    float *d_packed = nullptr;      // weight images
    size_t packed_cap = 0;
    uint4 *h_rows = nullptr;        // mapped
    unsigned *h_symw = nullptr;     // mapped
    unsigned *h_ctl = nullptr;      // mapped
    size_t rows_cap_total = 0;
    bool attr_set[2] = {false, false};  // per kernel form
    unsigned epoch = 1;
    int max_smem[2] = {0, 0};
};
std::mutex g_states_mu;
FlowState *g_states[64] = {nullptr};

FlowState *flow_state()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_states_mu);
    if (!g_states[dev]) g_states[dev] = new FlowState();
    return g_states[dev];
}

std::atomic<int> g_flow_threads{0};   // host coder threads per call (0 = automatic)

}  // namespace

// Host coder threads for a call that codes nimg bitstreams: explicit setting (pcx_flow_set_threads / PCX_CODER_THREADS) or the
// cores this process may use divided by the local ranks (torchrun exports LOCAL_WORLD_SIZE; one process per GPU shares the
// host), minus one for the Python thread; never more than one per image.
int pcx_host_coder_threads(int nimg)
{
    int t = g_flow_threads.load();
    if (t <= 0) {
        if (const char *e = getenv("PCX_CODER_THREADS")) t = atoi(e);
    }
    if (t <= 0) {
        int cores = (int)std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        int local = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) local = atoi(e) > 0 ? atoi(e) : 1;
        t = cores / local - 1;
    }
    return t < 1 ? 1 : (t > nimg ? nimg : t);
}

namespace {
inline int flow_host_threads(int nimg) { return pcx_host_coder_threads(nimg); }

}  // namespace

int pcx_flow_set_threads(int n) { return g_flow_threads.exchange(n); }

int pcx_wave_decode_flow(const pcx_wave_net &n, pcx_coder *const *coders, long long *n_symbols, cudaStream_t s, bool *unsupported)
{
    *unsupported = false;
    const int Hf = n.h * n.npart, nsteps = pcx_wave_steps(&n), nrep = n.nb * n.nimg, nplanes = Hf + n.W - 1;
    FlowState *stp = flow_state();
    PCX_REQUIRE(stp != nullptr, "no current CUDA device");
    FlowState &st = *stp;
    std::lock_guard<std::mutex> lock(st.mu);

    // ---- which form of the kernel: tasks of one layer of the widest step per warp of the 256 x 2 form
    int widest = 1;
    for (int step = 0; step < nsteps; step++) {
        const int p0 = step - n.G + 1 < 0 ? 0 : step - n.G + 1, p1 = step < nplanes - 1 ? step + 1 : nplanes;
        if (p1 > p0) widest = std::max(widest, n.h_start[p1] - n.h_start[p0]);
    }
    const long tasks_per_layer = (long)widest * n.nb * n.nimg;
    int form = 2 * tasks_per_layer >= 3l * pcx_sm_count() * 16 ? 1 : 0;       // 1.5 or more tasks per warp and layer: small blocks
    if (const char *e = getenv("PCX_FLOW_FORM")) form = atoi(e) == 1 ? 1 : 0;
    const int threads = form == 1 ? 128 : 256;
    const size_t smem = sizeof(float) * 2 * (4 * (size_t)n.G * 3 * 25 + 8);
    const void *kernel = form == 1 ? (const void *)wave_flow_kernel<128, 4, 32> : (const void *)wave_flow_kernel<256, 2, 64>;
    if (!st.attr_set[form]) {
        int dev = 0, lim = 0;
        PCX_CUDA(cudaGetDevice(&dev));
        PCX_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cudaFuncAttributes fa;
        PCX_CUDA(cudaFuncGetAttributes(&fa, kernel));
        st.max_smem[form] = lim - (int)fa.sharedSizeBytes - 1024;            // what is left for dynamic shared memory
        PCX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, st.max_smem[form]));
        st.attr_set[form] = true;
    }
    int per_sm = 0;
    if (smem > (size_t)st.max_smem[form]) { *unsupported = true; return PCX_OK; }
    PCX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    const int max_grid = pcx_sm_count() * per_sm;
    int B = n.nimg > 0 ? max_grid / n.nimg : 0;                  // blocks per image
    if (const char *e = getenv("PCX_FLOW_BLOCKS")) B = atoi(e) >= 2 && atoi(e) < B ? atoi(e) : B;      // debugging / tuning
    const int smax = getenv("PCX_FLOW_SMAX") ? atoi(getenv("PCX_FLOW_SMAX")) : 0;
    if (per_sm < 1 || B < 2 || nsteps >= 32767) { *unsupported = true; return PCX_OK; }

    // ---- per-step schedule (per image): runs of at most S cells per (net, plane); block i of the image takes runs i, i + B, ...
    int maxcount = 1;
    std::vector<int4> sched(nsteps);
    std::vector<int> counts(nsteps);
    const int wpb = threads / 32;
    for (int step = 0; step < nsteps; step++) {
        int p0 = step - n.G + 1 < 0 ? 0 : step - n.G + 1;
        int p1 = step < nplanes - 1 ? step + 1 : nplanes;         // planes [p0, p1) (entropy_conv_cuda_v2.cu:389-391)
        if (p0 > p1) p0 = p1;
        counts[step] = n.h_start[p1] - n.h_start[p0];
        maxcount = std::max(maxcount, counts[step]);
        int maxcells = 1;
        for (int q = p0; q < p1; q++) maxcells = std::max(maxcells, n.h_start[q + 1] - n.h_start[q]);
        int S = wpb, nchunk = 0;
        long best = -1;
        for (int cand = wpb; cand < maxcells + wpb && (smax <= 0 || cand <= smax || cand == wpb); cand += wpb) {
            int chunks = 0;
            for (int q = p0; q < p1; q++) chunks += n.nb * ceil_div(n.h_start[q + 1] - n.h_start[q], cand);
            const long cost = (long)ceil_div(chunks, B) * (cand / wpb);
            if (best < 0 || cost <= best) { best = cost; S = cand; nchunk = chunks; }      // ties: the longer run (fewer weight stagings)
        }
        sched[step] = make_int4(p0, p1 - p0, S, nchunk);
    }
    const int rows_cap = (maxcount + 63) / 64 * 64;

    // ---- buffers
    FlowNet f;
    memset(&f, 0, sizeof(f));
    StepNet &d = f.net;
    d.nlayers = n.nlayers; d.nb = n.nb; d.nimg = n.nimg; d.npart = n.npart; d.G = n.G; d.h = n.h; d.W = n.W; d.pad = n.pad;
    d.nstep = n.nstep; d.ng = n.ng;
    d.gmm_bias = n.gmm_bias; d.gmm_total = n.gmm_total; d.gmm_beta = n.gmm_beta; d.input_bias = n.input_bias;
    PCX_REQUIRE(make_bands(d.bands, n.wl, n.npart) == 0, "bad band description");
    d.hband = n.d_band; d.hrow = n.d_row; d.hcol = n.d_col; d.htw = n.d_tw; d.order = n.d_order;
    d.prev = nullptr; d.cdf = nullptr; d.bar = nullptr; d.dbg = nullptr;
    const int G8 = (n.G + 7) / 8 * 8;
    const i64 planes = (i64)nrep * n.npart;
    const i64 in_elems = planes * (n.h + 2 * n.pad) * (n.W + 2 * n.pad);
    std::vector<i64> off(n.nlayers + 2);
    off[0] = 0;
    off[1] = in_elems * G8 * n.layers[0].gi;
    for (int L = 0; L < n.nlayers; L++)
        off[L + 2] = off[L + 1] + planes * (n.h + 2 * n.layers[L].pad_out) * (n.W + 2 * n.layers[L].pad_out) * G8 * 3;
    PCX_REQUIRE(off[n.nlayers + 1] < 0x7fffffffll * 8, "scratch too large");
    for (int L = 0; L <= n.nlayers; L++)
        PCX_REQUIRE((off[L + 1] - off[L]) / (L == 0 ? G8 * n.layers[0].gi : G8 * 3) < 0x7fffffffll, "layer %d: cell offsets exceed 31 bits", L);
    const int ncell_total = n.h_start[nplanes];                                   // valid cells of one image
    const size_t cell_off = (sizeof(float) * (size_t)off[n.nlayers + 1] + 255) / 256 * 256;
    const size_t need_bytes = cell_off + sizeof(int4) * (size_t)ncell_total + 256;
    if (need_bytes > st.scratch_bytes) {
        if (st.scratch) cudaFree(st.scratch);
        st.scratch = nullptr;
        st.scratch_bytes = 0;
        PCX_CUDA(cudaMalloc((void **)&st.scratch, need_bytes));
        st.scratch_bytes = need_bytes;
    }
    if (!st.dev_ctl) PCX_CUDA(cudaMalloc((void **)&st.dev_ctl, sizeof(unsigned) * (16 + 1024)));
    PCX_REQUIRE(n.nimg <= 1024, "at most 1024 images per call");
    if ((size_t)(nplanes + 1) > st.start_cap) {
        if (st.d_start) cudaFree(st.d_start);
        st.d_start = nullptr; st.start_cap = 0;
        PCX_CUDA(cudaMalloc((void **)&st.d_start, sizeof(int) * (size_t)(nplanes + 1)));
        st.start_cap = (size_t)(nplanes + 1);
    }
    if ((size_t)nsteps > st.sched_cap) {
        if (st.d_sched) cudaFree(st.d_sched);
        st.d_sched = nullptr; st.sched_cap = 0;
        PCX_CUDA(cudaMalloc((void **)&st.d_sched, sizeof(int4) * (size_t)nsteps));
        st.sched_cap = (size_t)nsteps;
    }
    const size_t rows_total = (size_t)rows_cap * n.nimg;
    if (rows_total > st.rows_cap_total) {
        if (st.h_rows) cudaFreeHost(st.h_rows);
        if (st.h_symw) cudaFreeHost(st.h_symw);
        st.h_rows = nullptr; st.h_symw = nullptr; st.rows_cap_total = 0;
        PCX_CUDA(cudaHostAlloc((void **)&st.h_rows, rows_total * sizeof(uint4), cudaHostAllocMapped | cudaHostAllocPortable));
        PCX_CUDA(cudaHostAlloc((void **)&st.h_symw, rows_total * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable));
        st.rows_cap_total = rows_total;
    }
    if (!st.h_ctl) PCX_CUDA(cudaHostAlloc((void **)&st.h_ctl, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(st.h_rows, 0, rows_total * sizeof(uint4));             // tag 0 = never written
    memset(st.h_symw, 0, rows_total * sizeof(unsigned));
    memset(st.h_ctl, 0, 64);
    PCX_CUDA(cudaHostGetDevicePointer((void **)&f.rows, st.h_rows, 0));
    PCX_CUDA(cudaHostGetDevicePointer((void **)&f.symw, st.h_symw, 0));
    PCX_CUDA(cudaHostGetDevicePointer((void **)&f.host_ctl, st.h_ctl, 0));

    {
        const size_t nwords = cell_off / sizeof(unsigned);
        fill_u32_kernel<<<pcx_sm_count() * 8, 256, 0, s>>>(reinterpret_cast<unsigned *>(st.scratch), nwords, FLOW_SENTINEL);
        PCX_LAUNCHED();
    }
    PCX_CUDA(cudaMemsetAsync(st.dev_ctl, 0, sizeof(unsigned) * (16 + 1024), s));
    const int wbuf = 4 * n.G * 3 * 25 + 8;
    const size_t packed_floats = (size_t)n.nlayers * n.nb * n.G * wbuf;
    if (packed_floats > st.packed_cap) {
        if (st.d_packed) cudaFree(st.d_packed);
        st.d_packed = nullptr; st.packed_cap = 0;
        PCX_CUDA(cudaMalloc((void **)&st.d_packed, sizeof(float) * packed_floats));
        st.packed_cap = packed_floats;
    }
    PCX_CUDA(cudaMemcpyAsync(st.d_start, n.h_start, sizeof(int) * (size_t)(nplanes + 1), cudaMemcpyHostToDevice, s));
    PCX_CUDA(cudaMemcpyAsync(st.d_sched, sched.data(), sizeof(int4) * (size_t)nsteps, cudaMemcpyHostToDevice, s));
    for (int L = 0; L < n.nlayers; L++) {
        const pcx_wave_layer &l = n.layers[L];
        StepLayer &sl = d.L[L];
        sl.weight = l.weight; sl.bias = l.bias; sl.act = l.act;
        sl.in = st.scratch + off[L];
        sl.out = st.scratch + off[L + 1];
        sl.add = nullptr;
        if (l.add) {
            for (int M = 0; M < L; M++)
                if (n.layers[M].out == l.add) sl.add = st.scratch + off[M + 1];
            PCX_REQUIRE(sl.add != nullptr, "layer %d: the residual source must be the output of an earlier layer", L);
            PCX_REQUIRE(n.layers[L].pad_out == n.pad, "layer %d: residual add on an unpadded output", L);
        }
        sl.gi = l.gi; sl.pad_out = l.pad_out; sl.constrain = l.constrain;
        sl.cp_in = G8 * l.gi; sl.cp_out = G8 * 3;
        PCX_REQUIRE(L == 0 || (l.in == n.layers[L - 1].out && l.gi == 3 && n.layers[L - 1].pad_out == n.pad),
                    "layer %d must read the padded output of layer %d", L, L - 1);
    }
    d.sym_nchw = n.layers[0].in;
    int4 *d_cell = reinterpret_cast<int4 *>(reinterpret_cast<char *>(st.scratch) + cell_off);
    step_cellinfo_kernel<<<ceil_div(ncell_total, 256), 256, 0, s>>>(n.d_order, d_cell, ncell_total, n.h, n.W);
    PCX_LAUNCHED();
    d.cell = d_cell;
    {
        float fo = (float)(1 - 1e-6);
        if ((double)fo < 1 - 1e-6) fo = nextafterf(fo, 2.0f);
        d.halo_one = fo;
    }
    PCX_CUDA(cudaMemsetAsync(n.layers[0].in, 0, sizeof(float) * (size_t)nrep * n.npart * n.G * (n.h + 2 * n.pad) * (n.W + 2 * n.pad), s));
    flow_pack_weights_kernel<<<pcx_sm_count() * 4, 256, 0, s>>>(d, st.d_packed, wbuf);      // after d.L[] is filled in above
    PCX_LAUNCHED();
    f.packed = st.d_packed;
    f.nsteps = nsteps; f.rows_cap = rows_cap; f.blocks_per_img = B;
    const unsigned salt = (st.epoch++ * 9973u) & 0x7fffu;
    f.tag_salt = salt;
    f.start = st.d_start; f.sched = st.d_sched; f.dev_ctl = st.dev_ctl;
    f.timeout_ns = 20ull * 1000000000ull;
    if (const char *e = getenv("PCX_FLOW_TIMEOUT_MS")) f.timeout_ns = (unsigned long long)atoll(e) * 1000000ull;

#ifdef PCX_FLOW_PROBE
    unsigned long long *d_probe = nullptr;
    PCX_CUDA(cudaMalloc((void **)&d_probe, 64));
    PCX_CUDA(cudaMemsetAsync(d_probe, 0, 64, s));
    d.dbg = d_probe;
#endif
    const char *trace_path = getenv("PCX_WAVE_TRACE");
    unsigned long long *h_trace = nullptr;
    if (trace_path) {
        if (cudaHostAlloc((void **)&h_trace, sizeof(unsigned long long) * 16 * (size_t)(nsteps + 1), cudaHostAllocMapped) == cudaSuccess) {
            memset(h_trace, 0, sizeof(unsigned long long) * 16 * (size_t)(nsteps + 1));
            if (cudaHostGetDevicePointer((void **)&f.trace, h_trace, 0) != cudaSuccess) f.trace = nullptr;
        } else {
            (void)cudaGetLastError();
            h_trace = nullptr;
        }
    }

    // The launch goes through a helper thread: the kernel needs this thread's decoder loop below to make progress, and a
    // launch call that only returns when the kernel has finished (a profiler that serialises launches, e.g. the ncu launch-list
    // pass) would otherwise dead-lock against it until the device time-out.  Normally the call returns at once.
    void *args[] = {(void *)&f};
    int launch_dev = 0;
    PCX_CUDA(cudaGetDevice(&launch_dev));
    std::atomic<int> launch_state{0};                  // 0 pending, 1 returned ok, -1 failed
    cudaError_t le = cudaSuccess;
    const int grid = B * n.nimg;
    std::thread launcher([&] {
        cudaError_t e = cudaSetDevice(launch_dev);
        if (e == cudaSuccess) e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(threads), args, smem, s);
        le = e;
        launch_state.store(e == cudaSuccess ? 1 : -1, std::memory_order_release);
    });
    g_pcx_launches.fetch_add(1, std::memory_order_relaxed);

    // ---- host decoders: T threads, thread t serves images t, t + T, ...; each image is its own state machine
    const int T = getenv("PCX_WAVE_DUMP") ? 1 : flow_host_threads(n.nimg);
    std::atomic<int> status{PCX_OK};
    std::atomic<long long> total{0};
    std::mutex msg_mu;
    std::string message;
    std::vector<uint4> log_rows;
    std::vector<int4> log_meta;
    FILE *dump = nullptr;                              // PCX_WAVE_DUMP=<file>: every decoded row (debugging; one host thread)
    if (const char *e = getenv("PCX_WAVE_DUMP")) dump = fopen(e, "w");
    volatile unsigned *h_ctl = st.h_ctl;
    auto serve = [&](int t0) {
        struct Img { int step, pos; bool done; };
        std::vector<Img> im;
        std::vector<int> ids;
        std::vector<uint4> snap;
        for (int i = t0; i < n.nimg; i += T) { ids.push_back(i); im.push_back({0, 0, false}); }
        size_t left = ids.size();
        long long mine = 0;
        unsigned idle = 0;
        auto last_progress = std::chrono::steady_clock::now();
        while (left > 0 && status.load(std::memory_order_relaxed) == PCX_OK) {
            bool progressed = false;
            for (size_t q = 0; q < ids.size(); q++) {
                Img &m = im[q];
                if (m.done) continue;
                while (m.step < nsteps && counts[m.step] == 0) m.step++;
                if (m.step >= nsteps) { m.done = true; left--; progressed = true; continue; }
                const int cnt = counts[m.step], i = ids[q];
                const size_t base = (size_t)i * rows_cap + m.pos;
                int got = 0;
                const uint16_t *src = reinterpret_cast<const uint16_t *>(st.h_rows + base);
                if (dump) {                            // decode from a snapshot: the device reuses a row as soon as the step is decoded
                    snap.resize((size_t)(cnt - m.pos));
                    memcpy(snap.data(), st.h_rows + base, sizeof(uint4) * (size_t)(cnt - m.pos));
                    src = reinterpret_cast<const uint16_t *>(snap.data());
                }
                const int rc = pcx_coder_decodes_rows16(coders[i], src, cnt - m.pos, 0x8000u | ((salt + (unsigned)m.step) & 0x7fffu),
                                                        0x800000u | ((salt & 0x7fu) << 16) | ((unsigned)(m.step + 1) & 0xffffu),
                                                        st.h_symw + base, &got);
                if (dump && got > 0) {                  // kept in memory, written after the decode: the host must stay fast
                    for (int r = 0; r < got; r++) {
                        uint4 rec = snap[(size_t)r];
                        log_rows.push_back(rec);
                        log_meta.push_back(make_int4(i, m.step, m.pos + r, (int)(st.h_symw[base + r] & 0xffu)));
                    }
                }
                if (rc < 0) {
                    // the message lives in this thread's error slot: hand it to the calling thread
                    pcx_append_error(" [image %d, wavefront step %d, row %d of %d]", i, m.step, m.pos + got, cnt);
                    std::lock_guard<std::mutex> lk(msg_mu);
                    if (status.load() == PCX_OK) { message = pcx_last_error(); status.store(rc); }
                    break;
                }
                if (got > 0) {
                    progressed = true;
                    m.pos += got;
                    if (m.pos == cnt) { mine += cnt; m.step++; m.pos = 0; }
                }
            }
            if (progressed) {
                idle = 0;
                continue;
            }
            _mm_pause();
            if ((++idle & 1023u) == 0) {
                if (h_ctl[1] != 0) { status.store(PCX_ECUDA); break; }                 // the device gave up (time-out)
                if (launch_state.load(std::memory_order_acquire) < 0) { status.store(PCX_ECUDA); break; }      // the kernel never started
                const auto now = std::chrono::steady_clock::now();
                if (idle == 1024u) last_progress = now;
                else if (std::chrono::duration<double>(now - last_progress).count() > 30.0) { status.store(PCX_ECUDA); break; }
                if (idle > (1u << 16)) sched_yield();                                   // oversubscribed host: give the core away
            }
        }
        total.fetch_add(mine);
    };
    std::vector<std::thread> workers;
    for (int t = 1; t < T; t++) workers.emplace_back(serve, t);
    serve(0);
    for (auto &w : workers) w.join();
    if (status.load() != PCX_OK) h_ctl[0] = 1;         // tell the kernel to stop waiting (before joining a launch call that may block on it)
    launcher.join();
    if (le != cudaSuccess) {
        if (h_trace) cudaFreeHost(h_trace);
        if (dump) fclose(dump);
        pcx_set_error("%s:%d cooperative launch of the decoder kernel -> %s", __FILE__, __LINE__, cudaGetErrorString(le));
        return PCX_ECUDA;
    }
    if (dump) {
        for (size_t q = 0; q < log_rows.size(); q++) {
            const uint16_t *row = reinterpret_cast<const uint16_t *>(&log_rows[q]);
            fprintf(dump, "%d %d %d |", log_meta[q].x, log_meta[q].y, log_meta[q].z);
            for (int j = 0; j < 7; j++) fprintf(dump, " %u", (unsigned)row[j]);
            fprintf(dump, " | %d\n", log_meta[q].w);
        }
        fclose(dump);
    }
    const int rc = status.load();
    if (rc != PCX_OK) h_ctl[0] = 1;                       // tell the kernel to stop waiting
    cudaError_t se = cudaStreamSynchronize(s);
#ifdef PCX_FLOW_PROBE
    {
        unsigned long long hp[8] = {0};
        cudaMemcpy(hp, d_probe, 64, cudaMemcpyDeviceToHost);
        cudaFree(d_probe);
        if (hp[5]) fprintf(stderr, "[flow probe] %llu tasks, cycles per task: setup + issue %.0f | prefix (load latency + FMAs) %.0f | wait for last elements %.0f | last FMAs %.0f | fold + store %.0f\n",
                           hp[5], (double)hp[0] / hp[5], (double)hp[1] / hp[5], (double)hp[2] / hp[5], (double)hp[3] / hp[5], (double)hp[4] / hp[5]);
    }
#endif
    const unsigned dev_err = h_ctl[1];
    if (h_trace) {
        if (FILE *fp = fopen(trace_path, "w")) {
            fprintf(fp, "# step rows | image 0 leader block, ns: tables + wait for symbols | after each layer | gmm rows | until next step\n");
            for (int stp_ = 0; stp_ < nsteps; stp_++) {
                const unsigned long long *t = h_trace + (size_t)stp_ * 16;
                fprintf(fp, "%d %d | %lld |", stp_, counts[stp_], (long long)(t[1] - t[0]));
                unsigned long long prev = t[1];
                for (int L = 0; L < 12 && L < n.nlayers; L++) {
                    fprintf(fp, " %lld", (long long)(t[2 + L] - prev));
                    prev = t[2 + L];
                }
                fprintf(fp, " | %lld %lld | layers %lld\n", (long long)(t[15] - t[14]), (long long)(t[16] - t[15]), (long long)(t[14] - t[1]));
            }
            fclose(fp);
        }
        cudaFreeHost(h_trace);
    }
    if (se != cudaSuccess) { pcx_set_error("%s:%d decoder kernel -> %s", __FILE__, __LINE__, cudaGetErrorString(se)); return PCX_ECUDA; }
    if (rc == PCX_ECUDA || (rc == PCX_OK && dev_err != 0)) {
        pcx_set_error("decoder kernel gave up waiting (code %u): host decoder stalled or device time-out", dev_err);
        return PCX_ECUDA;
    }
    if (rc != PCX_OK) {
        if (!message.empty()) pcx_set_error("%s", message.c_str());
        return rc;
    }
    if (n_symbols) *n_symbols = total.load();
    return PCX_OK;
}
