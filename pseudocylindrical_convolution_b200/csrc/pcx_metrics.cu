// pcx_metrics.cu - the evaluation side of `pseudo_codec.py --test`: the 14-viewport rectilinear projector
// (ProjectsOp / MultiProject, extension/projects_cuda.cu) and the Gaussian-window SSIM / MSE reductions
// (PCONV_operator/pytorch_ssim.py, pseudo_codec.py:270-284).  Metric code: float arithmetic follows the reference's
// expressions; results are compared with the reference within a tolerance (the reference's own SSIM runs through cuDNN).
#include "pcx_common.cuh"
#include <math.h>

namespace {

constexpr int NVIEW = 14;

// Rodrigues rotation matrices of 14 axis-angle vectors (projects_mrod, projects_cuda.cu:20-49); host, float like the reference
void rodrigues14(const float *x, const float *y, const float *z, float *data)
{
    for (int i = 0; i < NVIEW; i++) {
        float *m = data + i * 9;
        for (int k = 0; k < 9; k++) m[k] = 0.f;
        const float norm = sqrt(x[i] * x[i] + y[i] * y[i] + z[i] * z[i]);
        if (norm == 0) {
            m[0] = m[4] = m[8] = 1.f;
            continue;
        }
        const float tx = x[i] / norm, ty = y[i] / norm, tz = z[i] / norm;
        const float c = cos(norm), s = sin(norm);
        m[0] = c + (1 - c) * tx * tx;
        m[1] = (1 - c) * tx * ty - s * tz;
        m[2] = (1 - c) * tx * tz + s * ty;
        m[3] = (1 - c) * ty * tx + s * tz;
        m[4] = c + (1 - c) * ty * ty;
        m[5] = (1 - c) * ty * tz - s * tx;
        m[6] = (1 - c) * tz * tx - s * ty;
        m[7] = (1 - c) * tz * ty + s * tx;
        m[8] = c + (1 - c) * tz * tz;
    }
}

struct Rot14 { float r[NVIEW * 9]; };

// unit view rays of the h_out x w_out image plane (projects_init_xyz_kernel :7-19), rotated into each viewport
// (gmm_transpose_kernel :81-95), converted to ERP pixel coordinates (projects_cal_xyz_kernel :50-66)
__global__ void project_table_kernel(Rot14 rot, float *__restrict__ tf, int h_out, int w_out, float w_stride, float h_stride,
                                     float c_x, float c_y, float hx, float hy, float pi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int inner = h_out * w_out;
    if (i >= NVIEW * inner) return;
    const int v = i / inner, w = i % w_out, h = (i / w_out) % h_out;
    const float x = 1.;
    const float y = (w - c_x) * w_stride;
    const float z = (h - c_y) * h_stride;
    const float r = sqrt(x * x + y * y + z * z);
    const float xa = x / r, xb = y / r, xc = -z / r;
    const float *m = rot.r + v * 9;
    const float px = xa * m[0] + xb * m[1] + xc * m[2];
    const float py = xa * m[3] + xb * m[4] + xc * m[5];
    const float pz = xa * m[6] + xb * m[7] + xc * m[8];
    const float lat = asin(pz);
    float theta = atan(py / px);
    if (px <= 0) {
        if (py > 0) theta = theta + pi;
        else theta = theta - pi;
    }
    tf[(i64)i * 2] = theta / pi * hx + hx;
    tf[(i64)i * 2 + 1] = -2 * lat / pi * hy + hy;
}

// projects_forward_kernel / _nearest (:181-213): output index (viewport, image*channel, pixel)
__global__ void project_fwd_kernel(const float *__restrict__ in, const float *__restrict__ tf, float *__restrict__ out, i64 total,
                                   int inner, int hs, int ws, int planes, int nearest)
{
    for (i64 index = (i64)blockIdx.x * blockDim.x + threadIdx.x; index < total; index += (i64)gridDim.x * blockDim.x) {
        const int ps = (int)(index % inner);
        const i64 tn = (index / inner) % planes;
        const int tb = (int)(index / inner / planes);
        const float fx = tf[((i64)tb * inner + ps) * 2], fy = tf[((i64)tb * inner + ps) * 2 + 1];
        const float *img = in + tn * hs * (i64)ws;
        if (nearest) {
            const int tw = static_cast<int>(floor(fx + 0.5)) % ws;
            int th = static_cast<int>(floor(fy + 0.5));
            th = th >= hs ? hs - 1 : th;
            out[index] = img[(i64)th * ws + tw];
        } else {
            const int tw = static_cast<int>(floor(fx));
            const int th = static_cast<int>(floor(fy));
            const int pw = (tw + 1) % ws;
            const int ph = th + 1 >= hs ? hs - 1 : th + 1;
            const float tx = fx - tw;
            const float ty = fy - th;
            const float ntx = 1. - tx;
            const float nty = 1. - ty;
            out[index] = img[(i64)th * ws + tw] * ntx * nty + img[(i64)th * ws + pw] * tx * nty + img[(i64)ph * ws + tw] * ntx * ty +
                         img[(i64)ph * ws + pw] * tx * ty;
        }
    }
}

struct Gauss { float g[32]; };

// _ssim (pytorch_ssim.py:17-37): five zero-padded Gaussian-window correlations per pixel, then the SSIM map; each block leaves
// the sum of its pixels in `partial` (fixed order: deterministic)
__global__ void __launch_bounds__(256) ssim_kernel(const float *__restrict__ a, const float *__restrict__ b, Gauss win, int ws,
                                                   i64 planes, int h, int w, float *__restrict__ map, double *__restrict__ partial)
{
    __shared__ double red[256];
    const i64 total = planes * h * w;
    const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    double mine = 0.0;
    if (idx < total) {
        const int x = (int)(idx % w), y = (int)((idx / w) % h);
        const i64 p = idx / w / h;
        const float *pa = a + p * h * (i64)w, *pb = b + p * h * (i64)w;
        const int half = ws / 2;
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
        for (int j = 0; j < ws; j++) {
            const int yy = y + j - half;
            if (yy < 0 || yy >= h) continue;
            for (int i = 0; i < ws; i++) {
                const int xx = x + i - half;
                if (xx < 0 || xx >= w) continue;
                const float wt = win.g[j] * win.g[i];
                const float va = pa[(i64)yy * w + xx], vb = pb[(i64)yy * w + xx];
                m1 += wt * va;
                m2 += wt * vb;
                s11 += wt * (va * va);
                s22 += wt * (vb * vb);
                s12 += wt * (va * vb);
            }
        }
        const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu1_mu2 = m1 * m2;
        const float sigma1_sq = s11 - mu1_sq, sigma2_sq = s22 - mu2_sq, sigma12 = s12 - mu1_mu2;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float v = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2));
        if (map) map[idx] = v;
        mine = v;
    }
    red[threadIdx.x] = mine;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// mean of (a - b)^2 (b != nullptr) or of a: per-block partial sums in a fixed order
__global__ void __launch_bounds__(256) sqdiff_kernel(const float *__restrict__ a, const float *__restrict__ b, i64 total,
                                                     double *__restrict__ partial)
{
    __shared__ double red[256];
    double mine = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const float d = b ? a[i] - b[i] : a[i];
        mine += b ? (double)(d * d) : (double)d;
    }
    red[threadIdx.x] = mine;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void final_sum_kernel(const double *__restrict__ partial, int n, double scale, double *__restrict__ out)
{
    __shared__ double red[256];
    double mine = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine += partial[i];
    red[threadIdx.x] = mine;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0] * scale;
}

}  // namespace

extern "C" {

int pcx_project_table(const float *theta, const float *phi, float fov, int h_out, int w_out, int H, int W, float *d_tf, void *stream)
{
    PCX_REQUIRE(theta && phi && d_tf, "null pointer");
    PCX_REQUIRE(h_out > 1 && w_out > 1 && H > 0 && W > 0, "bad projector geometry");
    // projects_opt ctor (projects.hpp:8-19): angles arrive in units of pi
    const float pi = acos(-1.0);
    float th[NVIEW], ph[NVIEW];
    for (int i = 0; i < NVIEW; i++) { th[i] = theta[i] * pi; ph[i] = phi[i] * pi; }
    const float fovr = fov * pi;
    // projects_opt::init (:96-139)
    const float hfov = fovr * h_out / w_out / 2;
    const float wfov = fovr / 2;
    const float c_x = (w_out - 1) / 2.0;
    const float c_y = (h_out - 1) / 2.0;
    const float pi_2 = pi / 2;
    const float wangle = pi_2 - wfov, hangle = pi_2 - hfov;
    const float w_stride = 2 * sin(wfov) / sin(wangle) / (w_out - 1);
    const float h_stride = 2 * sin(hfov) / sin(hangle) / (h_out - 1);
    float r1[NVIEW * 9], r2[NVIEW * 9], xa[NVIEW], ya[NVIEW], za[NVIEW];
    for (int i = 0; i < NVIEW; i++) { xa[i] = 0; ya[i] = 0; za[i] = th[i]; }
    rodrigues14(xa, ya, za, r1);
    for (int i = 0; i < NVIEW; i++) {
        xa[i] = r1[i * 9 + 1] * (-ph[i]);
        ya[i] = r1[i * 9 + 4] * (-ph[i]);
        za[i] = r1[i * 9 + 7] * (-ph[i]);
    }
    rodrigues14(xa, ya, za, r2);
    Rot14 rot;
    for (int v = 0; v < NVIEW; v++)            // gmm_kernel (:67-80): r = r2 x r1
        for (int m = 0; m < 3; m++)
            for (int n = 0; n < 3; n++) {
                float sum = 0;
                for (int j = 0; j < 3; j++) sum += r2[v * 9 + m * 3 + j] * r1[v * 9 + j * 3 + n];
                rot.r[v * 9 + m * 3 + n] = sum;
            }
    // projects_opt::update (:140-152)
    const float hx = (W - 1) / 2.0, hy = (H - 1) / 2.0;
    const int count = NVIEW * h_out * w_out;
    project_table_kernel<<<ceil_div(count, 256), 256, 0, (cudaStream_t)stream>>>(rot, d_tf, h_out, w_out, w_stride, h_stride, c_x, c_y,
                                                                                 hx, hy, pi);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_project_fwd(const float *d_in, const float *d_tf, float *d_out, int N, int C, int H, int W, int h_out, int w_out, int nearest,
                    void *stream)
{
    PCX_REQUIRE(d_in && d_tf && d_out, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && h_out > 0 && w_out > 0, "bad shape");
    const i64 total = (i64)N * C * h_out * w_out * NVIEW;
    i64 want = (total + 255) / 256;
    const i64 cap = (i64)pcx_sm_count() * 16;
    project_fwd_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(d_in, d_tf, d_out, total, h_out * w_out, H, W,
                                                                                       N * C, nearest);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_ssim(const float *d_a, const float *d_b, long long planes, int h, int w, int window, float sigma, float *d_map,
             double *d_scratch, long long scratch_len, double *d_mean, void *stream)
{
    PCX_REQUIRE(d_a && d_b && d_scratch && d_mean, "null pointer");
    PCX_REQUIRE(planes > 0 && h > 0 && w > 0 && window >= 1 && window <= 31 && (window & 1), "bad SSIM geometry");
    const i64 total = planes * h * w;
    const i64 blocks = (total + 255) / 256;
    PCX_REQUIRE(blocks <= scratch_len && blocks < (1ll << 31), "SSIM scratch of %lld doubles is too small for %lld blocks", scratch_len, blocks);
    // gaussian(window, sigma) (pytorch_ssim.py:7-9): torch.Tensor of python floats, normalised in float32
    Gauss g;
    float sum = 0.f;
    for (int x = 0; x < window; x++) {
        g.g[x] = (float)exp(-(double)((x - window / 2) * (x - window / 2)) / (2.0 * sigma * sigma));
        sum += g.g[x];
    }
    for (int x = 0; x < window; x++) g.g[x] = g.g[x] / sum;
    cudaStream_t s = (cudaStream_t)stream;
    ssim_kernel<<<(int)blocks, 256, 0, s>>>(d_a, d_b, g, window, planes, h, w, d_map, d_scratch);
    PCX_LAUNCHED();
    final_sum_kernel<<<1, 256, 0, s>>>(d_scratch, (int)blocks, 1.0 / (double)total, d_mean);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_mean_sqdiff(const float *d_a, const float *d_b, long long total, double *d_scratch, long long scratch_len, double *d_mean,
                    void *stream)
{
    PCX_REQUIRE(d_a && d_scratch && d_mean && total > 0, "bad arguments");
    i64 blocks = (total + 255) / 256;
    const i64 cap = (i64)pcx_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    PCX_REQUIRE(blocks <= scratch_len, "scratch of %lld doubles is too small for %lld blocks", scratch_len, blocks);
    cudaStream_t s = (cudaStream_t)stream;
    sqdiff_kernel<<<(int)blocks, 256, 0, s>>>(d_a, d_b, total, d_scratch);
    PCX_LAUNCHED();
    final_sum_kernel<<<1, 256, 0, s>>>(d_scratch, (int)blocks, 1.0 / (double)total, d_mean);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
