// Classification CE + Hinton KD (KL) + top-1 accuracy of the student logits in ONE launch, with their gradients.
//
// Replaces, for the [B, n_cls] logits on either side of the MoMA loss in helper/loops_moma.py:278-279,350-355 :
//   loss_cls = nn.CrossEntropyLoss()(logit_s, labels)                                   (train_student_moma.py:294)
//   loss_div = DistillKL(T)(logit_s, logit_t) = KLDivLoss('batchmean')(log_softmax(s/T), softmax(t/T)) * T^2   (distiller_zoo/KD.py:7-17)
//   top-1    = accuracy(logit_s, labels, topk=(1,))[0]                                   (helper/util.py:71-85)
// The reference runs ~12 small kernels and two .item() syncs for these; here one 256-thread CTA produces the three
// device scalars and both gradients (d loss_cls / d s, d loss_div / d s), so the backward is a scalar multiply-add.
// n_cls <= 1024; one warp per row, rows summed in a fixed order (deterministic).
#include <math_constants.h>
#include "common.cuh"

namespace moma {

constexpr int kKdThreads = 256;

__global__ void __launch_bounds__(kKdThreads)
cls_kd_kernel(const float* __restrict__ ls, const float* __restrict__ lt, const int64_t* __restrict__ labels, int B, int C,
              float T, float* __restrict__ out /* loss_cls, loss_div, acc_pct */, float* __restrict__ g_cls,
              float* __restrict__ g_div) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float s_ce[kKdThreads / 32], s_kl[kKdThreads / 32], s_ok[kKdThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kKdThreads / 32;
    const float invT = 1.0f / T, invB = 1.0f / (float)B;
    float ce = 0.f, kl = 0.f, ok = 0.f;                      // per-warp partial sums over its rows (lane 0 holds them)
    for (int r = warp; r < B; r += nw) {
        const float* s = ls + (int64_t)r * C;
        const float* t = lt + (int64_t)r * C;
        const int lab = (int)labels[r];
        // maxima and arg-max of s (lowest index among exact ties; torch.topk leaves ties unspecified)
        float ms = -CUDART_INF_F, mt = -CUDART_INF_F;
        int am = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = s[c];
            if (v > ms) { ms = v; am = c; }
            mt = fmaxf(mt, t[c]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, ms, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (om > ms || (om == ms && oa < am)) { ms = om; am = oa; }
        }
        mt = warp_max(mt);
        float z1 = 0.f, zs = 0.f, zt = 0.f;                 // sum exp(s - ms), sum exp((s - ms)/T), sum exp((t - mt)/T)
        for (int c = lane; c < C; c += 32) {
            z1 += expf(s[c] - ms);
            zs += expf((s[c] - ms) * invT);
            zt += expf((t[c] - mt) * invT);
        }
        z1 = warp_sum(z1); zs = warp_sum(zs); zt = warp_sum(zt);
        const float l1 = logf(z1), lzs = logf(zs), lzt = logf(zt);
        float klr = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float p1 = expf(s[c] - ms - l1);                          // softmax(s)
            const float lps = (s[c] - ms) * invT - lzs;                     // log_softmax(s / T)
            const float lpt = (t[c] - mt) * invT - lzt;                     // log_softmax(t / T)
            const float pt = expf(lpt);
            if (pt > 0.f) klr += pt * (lpt - lps);                          // xlogy semantics of KLDivLoss: 0 * log 0 = 0
            g_cls[(int64_t)r * C + c] = (p1 - (c == lab ? 1.f : 0.f)) * invB;
            g_div[(int64_t)r * C + c] = (expf(lps) - pt) * T * invB;        // T^2 / B * d/ds KL = T / B * (softmax(s/T) - p_t)
        }
        klr = warp_sum(klr);
        if (lane == 0) {
            ce += (lab >= 0 && lab < C) ? (ms + l1 - s[lab]) : 0.f;
            kl += klr;
            ok += (am == lab) ? 1.f : 0.f;
        }
    }
    if (lane == 0) { s_ce[warp] = ce; s_kl[warp] = kl; s_ok[warp] = ok; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int w = 0; w < nw; ++w) { a += s_ce[w]; b += s_kl[w]; c += s_ok[w]; }
        out[0] = a * invB;
        out[1] = b * invB * T * T;
        out[2] = c * (100.0f * invB);
    }
}

}  // namespace moma

using namespace moma;

extern "C" __attribute__((visibility("default"))) int moma_cls_kd(const float* logit_s, const float* logit_t, const int64_t* labels,
                                                                 int64_t B, int64_t n_cls, float T, float* out3, float* grad_cls,
                                                                 float* grad_div, moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && n_cls > 0 && n_cls <= 1024 && B < (1ll << 24), MOMA_ERR_INVALID, "cls_kd: bad shape B=%lld n_cls=%lld",
                 (long long)B, (long long)n_cls);
    MOMA_REQUIRE(logit_s && logit_t && labels && out3 && grad_cls && grad_div, MOMA_ERR_INVALID, "cls_kd: null pointer");
    MOMA_REQUIRE(T > 0.f, MOMA_ERR_INVALID, "cls_kd: temperature must be positive");
    launch_pdl(cls_kd_kernel, dim3(1), dim3(kKdThreads), 0, as_stream(stream), logit_s, logit_t, labels, (int)B, (int)n_cls, T, out3,
               grad_cls, grad_div);
    MOMA_CUDA_LAUNCH_CHECK("cls_kd");
    note_launches(1);
    return MOMA_OK;
}
