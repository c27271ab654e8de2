// TEST INFRASTRUCTURE (built into tests/libanimnerf_b200_ref.so, never into the product library):
// fp32 SIMT restatement of the NeRF MLP forward: one thread per point, activations in
// (thread-interleaved) local memory, weights read through the warp-broadcast read-only path from
// the flat fp32 copy inside the packed buffer.  It is the on-device fp32 reference the tcgen05
// kernel is bisected against in the GPU tests (tests/util.py: ref_mlp()).  Follows
// models/embedding.py:22-39 and models/nerf.py:129-175.
#include "../../anim-nerf_b200/csrc/common.cuh"
#include "../../anim-nerf_b200/csrc/mlp_layout.cuh"

__global__ void __launch_bounds__(128)
mlp_fwd_ref_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ xyz_cano,
                   const int32_t* __restrict__ cidx, const int32_t* __restrict__ count, int64_t n_max,
                   float* __restrict__ sigma, float* __restrict__ rgb)
{
    using namespace mlp;
    const float* flat = (const float*)(packed + FLAT_OFF);
    int64_t n = n_max;
    if (cidx) { const int64_t c = *count; n = c < n_max ? c : n_max; }
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = cidx ? cidx[p] : p;
        float e[ENC], a[W], o[W];
        const float x[3] = {xyz_cano[id * 3], xyz_cano[id * 3 + 1], xyz_cano[id * 3 + 2]};
        for (int c = 0; c < 3; ++c) e[c] = x[c];
        for (int k = 0; k < 10; ++k) {
            const float f = (float)(1 << k);
            for (int c = 0; c < 3; ++c) { e[3 + 6 * k + c] = sinf(f * x[c]); e[3 + 6 * k + 3 + c] = cosf(f * x[c]); }
        }
        for (int l = 0; l < 8; ++l) {
            const float* Wl = flat + flat_w_off(l);
            const float* bl = flat + flat_b_off(l);
            const int in = lin_in(l);
            for (int r = 0; r < W; ++r) {
                float acc = __ldg(bl + r);
                const float* wr = Wl + (int64_t)r * in;
                if (l == 0) { for (int c = 0; c < ENC; ++c) acc += __ldg(wr + c) * e[c]; }
                else if (l == 4) {
                    for (int c = 0; c < ENC; ++c) acc += __ldg(wr + c) * e[c];
                    for (int c = 0; c < W; ++c) acc += __ldg(wr + ENC + c) * a[c];
                } else { for (int c = 0; c < W; ++c) acc += __ldg(wr + c) * a[c]; }
                o[r] = fmaxf(acc, 0.f);
            }
            for (int r = 0; r < W; ++r) a[r] = o[r];
        }
        {   // sigma head (raw)
            const float* ws = flat + flat_w_off(10);
            float acc = __ldg(flat + flat_b_off(10));
            for (int c = 0; c < W; ++c) acc += __ldg(ws + c) * a[c];
            sigma[id] = acc;
        }
        {   // final (no activation)
            const float* Wl = flat + flat_w_off(8);
            const float* bl = flat + flat_b_off(8);
            for (int r = 0; r < W; ++r) {
                float acc = __ldg(bl + r);
                for (int c = 0; c < W; ++c) acc += __ldg(Wl + (int64_t)r * W + c) * a[c];
                o[r] = acc;
            }
        }
        {   // dir (ReLU, 128) then rgb (sigmoid)
            const float* Wl = flat + flat_w_off(9);
            const float* bl = flat + flat_b_off(9);
            for (int r = 0; r < 128; ++r) {
                float acc = __ldg(bl + r);
                for (int c = 0; c < W; ++c) acc += __ldg(Wl + (int64_t)r * W + c) * o[c];
                a[r] = fmaxf(acc, 0.f);
            }
            const float* Wr = flat + flat_w_off(11);
            const float* br = flat + flat_b_off(11);
            for (int j = 0; j < 3; ++j) {
                float acc = __ldg(br + j);
                for (int c = 0; c < 128; ++c) acc += __ldg(Wr + j * 128 + c) * a[c];
                rgb[id * 3 + j] = 1.0f / (1.0f + expf(-acc));
            }
        }
    }
}

extern "C" int an_test_mlp_fwd_ref(const void* packed, const float* xyz_cano, const int32_t* cidx, const int32_t* count,
                                  int64_t n_max, float* sigma, float* rgb, void* stream)
{
    if (!packed || !xyz_cano || !sigma || !rgb || n_max <= 0) return AN_ERR_ARG;
    const int64_t want = (n_max + 127) / 128;
    const int64_t cap = (int64_t)an_num_sms() * 8;
    mlp_fwd_ref_kernel<<<(int)(want < cap ? want : cap), 128, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)packed, xyz_cano, cidx, count, n_max, sigma, rgb);
    AN_CHECK_LAUNCH();
    return AN_OK;
}
