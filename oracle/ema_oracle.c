/* C part of the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * moma_oracle_ema_f32 restates learning/contrast_trainer.py:207-211
 *     p2.data.mul_(m).add_(p1.detach().data, alpha=(1 - m))
 * in the reference's arithmetic type (fp32) with its two roundings:
 *     t  = fl32(p2 * m)                 -- ATen mul_ (scalar cast to float)
 *     p2 = fl32(fma(alpha, p1, t))      -- ATen add_(other, alpha): fused
 *                                          multiply-add (Vectorized::fmadd on
 *                                          CPU, FFMA on the GPU)
 * fmaf() gives the single-rounding FMA that numpy cannot express.
 * Pinned against the reference run on CPU by tests/test_oracle_golden.py.
 *
 * moma_oracle_enqueue_ids restates MoMA/mem_moco.py:25-26
 *     ids = fmod(arange(n) + index, K).long()
 */
#include <math.h>
#include <stdint.h>

void moma_oracle_ema_f32(float *p_ema, const float *p, int64_t n, float m, float alpha)
{
    for (int64_t i = 0; i < n; ++i) {
        float t = p_ema[i] * m;
        p_ema[i] = fmaf(alpha, p[i], t);
    }
}

void moma_oracle_enqueue_ids(int64_t *out, int64_t n, int64_t index, int64_t K)
{
    for (int64_t j = 0; j < n; ++j)
        out[j] = (j + index) % K;
}

/* moma_oracle_sgd_ema_f32 restates torch.optim.SGD(momentum, weight_decay; dampening 0, no Nesterov).step()
 * (train_student_moma.py:389-392, helper/loops_moma.py:361) followed by momentum_update
 * (learning/contrast_trainer.py:207-211) with the rounding sequence of the ATen kernels:
 *     d   = fl(fma(wd, p, g))               -- grad.add(param, alpha=weight_decay)
 *     buf = first ? d : fl(fl(buf * mu) + d) -- buf.mul_(momentum).add_(d, alpha=1)
 *     p   = fl(fma(-lr, buf, p))            -- param.add_(buf, alpha=-lr)
 *     ema = fl(fma(alpha, p, fl(ema * m)))
 * Pinned against torch.optim.SGD + the reference's momentum_update on CPU by tests/test_oracle_round2.py. */
void moma_oracle_sgd_ema_f32(float *p, const float *g, float *buf, float *ema, int64_t n, float lr, float mu, float wd,
                             int first, float m, float alpha)
{
    for (int64_t i = 0; i < n; ++i) {
        float d = fmaf(wd, p[i], g[i]);
        float b = first ? d : (buf[i] * mu) + d;
        buf[i] = b;
        p[i] = fmaf(-lr, b, p[i]);
        float t = ema[i] * m;
        ema[i] = fmaf(alpha, p[i], t);
    }
}
