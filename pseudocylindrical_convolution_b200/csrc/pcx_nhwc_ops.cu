// pcx_nhwc_ops.cu - the small channels-last operators that sit between the tensor-core convolutions of the analysis /
// synthesis transforms: in-place halo refresh (PseudoPadV2 on an already padded tile buffer), depth-to-space (Dtow d2w)
// into the interior of the next padded buffer, and the elementwise square that feeds the GDN contraction.
// All are HBM-bound, tiny next to the convolutions, and keep the reference's expression shapes (lerp2_ref).
#include "pcx_common.cuh"

namespace {

struct HaloParams {
    int C4, h, W, pad, rows, pitch, y0, x0;
    i64 planes;
};

// One thread per (ring cell, float4 of channels).  Ring cells of plane (image, band g): all cells of the padded tile
// [y0-pad, y0+h+pad) x [x0-pad, x0+wl+pad) that are not interior.  Value = PseudoPadV2 (pseudo_pad.cu:39-96): halo rows are
// the 2-tap interpolation of the neighbour band's row (pseudo_pad.cu:57-79), and every row - halo rows included - wraps
// in longitude modulo wl (pseudo_pad.cu:82-96).  Only interior cells are read and only ring cells are written.
__global__ void halo_fill_nhwc_kernel(float *__restrict__ buf, const int *__restrict__ hband, const int *__restrict__ hrow,
                                      const int *__restrict__ hcol, const float *__restrict__ htw, Bands bands, HaloParams P)
{
    const int npart = bands.npart;
    const int pad = P.pad, h = P.h, W = P.W;
    const i64 plane = blockIdx.y;
    const int g = (int)(plane % npart);
    const i64 img = plane / npart;
    const int wl = bands.wl[g];
    const int ringw = wl + 2 * pad;
    const int ncell = 2 * pad * ringw + 2 * pad * h;           // ring cells of one plane: 32-bit index arithmetic (the host checks)
    const i64 C = (i64)P.C4 * 4;
    const int total = ncell * P.C4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int cell = idx / P.C4;
        const int c4 = idx - cell * P.C4;
        int ty, tx;                                  // position inside the padded tile
        if (cell < 2 * pad * ringw) {                // halo rows (full ring width)
            const int rr = cell / ringw;
            tx = cell - rr * ringw;
            ty = rr < pad ? rr : h + rr;             // rr in [pad, 2 pad) -> rows h+pad .. h+2pad-1
        } else {                                     // left / right columns of the interior rows
            const int k = cell - 2 * pad * ringw;
            const int row = k / (2 * pad);
            ty = pad + row;
            const int j = k - row * (2 * pad);
            tx = j < pad ? j : wl + j;
        }
        // logical column with the longitude wrap
        int x = tx - pad;
        if (x < 0) x += wl;
        else if (x >= wl) x -= wl;
        float4 v;
        if (ty >= pad && ty < pad + h) {
            const float4 *src = reinterpret_cast<const float4 *>(buf + ((plane * P.rows + P.y0 + ty - pad) * (i64)P.pitch + P.x0 + x) * C) + c4;
            v = *src;
        } else {
            const int s = ty < pad ? 0 : 1, r = ty < pad ? ty : ty - pad - h;
            const int hr = (g * 2 + s) * pad + r;
            const int pg = hband[hr];
            const int srow = hrow[hr];
            const i64 e = (i64)hr * W + x;
            const int q = hcol[e];
            const int q1 = (q + 1 == bands.wl[pg]) ? 0 : q + 1;
            const float t = htw[e];
            const i64 sp = img * npart + pg;
            const float *rowp = buf + ((sp * P.rows + P.y0 + srow) * (i64)P.pitch + P.x0) * C;
            const float4 a = *(reinterpret_cast<const float4 *>(rowp + (i64)q * C) + c4);
            const float4 b = *(reinterpret_cast<const float4 *>(rowp + (i64)q1 * C) + c4);
            v = make_float4(lerp2_ref(a.x, b.x, t), lerp2_ref(a.y, b.y, t), lerp2_ref(a.z, b.z, t), lerp2_ref(a.w, b.w, t));
        }
        float4 *dst = reinterpret_cast<float4 *>(buf + ((plane * P.rows + P.y0 - pad + ty) * (i64)P.pitch + P.x0 - pad + tx) * C) + c4;
        *dst = v;
    }
}

struct DtowParams {
    int Co, h, W;                         // Co output channels, input extent h x W (output 2h x 2W)
    int in_rows, in_pitch, in_y0, in_x0, out_rows, out_pitch, out_y0, out_x0;
    i64 planes;
};

// dtow_forward_kernel (extension/dtow_cuda.cu:38-55), stride 2, channels last: input channel 4c + 2dy + dx of pixel (y,x)
// -> output channel c of pixel (2y+dy, 2x+dx).  One thread reads the float4 of one (pixel, c) and writes four pixels.
__global__ void dtow_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, DtowParams P)
{
    const i64 total = P.planes * P.h * P.W * P.Co;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(idx % P.Co);
        i64 t = idx / P.Co;
        const int x = (int)(t % P.W); t /= P.W;
        const int y = (int)(t % P.h);
        const i64 plane = t / P.h;
        const float4 v = *reinterpret_cast<const float4 *>(in + ((plane * P.in_rows + P.in_y0 + y) * (i64)P.in_pitch + P.in_x0 + x) * (4 * (i64)P.Co) + 4 * c);
        float *o = out + ((plane * P.out_rows + P.out_y0 + 2 * y) * (i64)P.out_pitch + P.out_x0 + 2 * x) * (i64)P.Co + c;
        o[0] = v.x;
        o[P.Co] = v.y;
        o += (i64)P.out_pitch * P.Co;
        o[0] = v.z;
        o[P.Co] = v.w;
    }
}

__global__ void square_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, i64 n4)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (i64)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        out[i] = make_float4(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w));
    }
}

}  // namespace

extern "C" {

int pcx_halo_fill_nhwc(float *d_buf, int N, int C, int h, int W, int npart, int pad, const int *wl, const int *d_band,
                       const int *d_row, const int *d_col, const float *d_tw, int rows, int pitch, int y0, int x0, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_buf && d_band && d_row && d_col && d_tw, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && C % 4 == 0 && h > 0 && W > 0 && pad > 0 && pad < 10, "bad shape N=%d C=%d h=%d W=%d pad=%d", N, C, h, W, pad);
    PCX_REQUIRE(y0 >= pad && x0 >= pad && y0 + h + pad <= rows && x0 + W + pad <= pitch, "padded tile does not fit the buffer planes");
    PCX_REQUIRE((reinterpret_cast<uintptr_t>(d_buf) & 15) == 0, "buffer must be 16-byte aligned");
    int wmax = 0;
    for (int i = 0; i < npart; i++) {
        PCX_REQUIRE(wl[i] >= 2 * pad && wl[i] <= W, "band %d width %d outside [2 pad, %d]", i, wl[i], W);
        if (wl[i] > wmax) wmax = wl[i];
    }
    HaloParams P;
    P.C4 = C / 4; P.h = h; P.W = W; P.pad = pad; P.rows = rows; P.pitch = pitch; P.y0 = y0; P.x0 = x0;
    P.planes = (i64)N * npart;
    PCX_REQUIRE(P.planes <= 65535, "too many planes");
    const i64 work = ((i64)2 * pad * (wmax + 2 * pad) + (i64)2 * pad * h) * P.C4;
    PCX_REQUIRE(work < (1ll << 30), "halo ring too large for 32-bit indexing");
    int bx = ceil_div(work, 256);
    if (bx > 1024) bx = 1024;
    halo_fill_nhwc_kernel<<<dim3(bx, (unsigned)P.planes), 256, 0, (cudaStream_t)stream>>>(d_buf, d_band, d_row, d_col, d_tw, b, P);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_dtow_nhwc(const float *d_in, float *d_out, int planes, int Co, int h, int W, int in_rows, int in_pitch, int in_y0,
                  int in_x0, int out_rows, int out_pitch, int out_y0, int out_x0, void *stream)
{
    PCX_REQUIRE(d_in && d_out && planes > 0 && Co > 0 && h > 0 && W > 0, "bad arguments");
    PCX_REQUIRE(in_y0 >= 0 && in_x0 >= 0 && in_y0 + h <= in_rows && in_x0 + W <= in_pitch, "input window outside the input plane");
    PCX_REQUIRE(out_y0 >= 0 && out_x0 >= 0 && out_y0 + 2 * h <= out_rows && out_x0 + 2 * W <= out_pitch, "output window outside the output plane");
    PCX_REQUIRE((reinterpret_cast<uintptr_t>(d_in) & 15) == 0, "input must be 16-byte aligned");
    DtowParams P;
    P.Co = Co; P.h = h; P.W = W;
    P.in_rows = in_rows; P.in_pitch = in_pitch; P.in_y0 = in_y0; P.in_x0 = in_x0;
    P.out_rows = out_rows; P.out_pitch = out_pitch; P.out_y0 = out_y0; P.out_x0 = out_x0;
    P.planes = planes;
    const i64 total = (i64)planes * h * W * Co;
    i64 blocks = (total + 255) / 256;
    const i64 cap = (i64)pcx_sm_count() * 16;
    dtow_nhwc_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, P);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_square(const float *d_in, float *d_out, long long n, void *stream)
{
    PCX_REQUIRE(d_in && d_out && n > 0 && n % 4 == 0, "bad arguments (n must be a multiple of 4)");
    PCX_REQUIRE(((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0, "buffers must be 16-byte aligned");
    const i64 n4 = n / 4;
    i64 blocks = (n4 + 255) / 256;
    const i64 cap = (i64)pcx_sm_count() * 16;
    square_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>((const float4 *)d_in, (float4 *)d_out, n4);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
