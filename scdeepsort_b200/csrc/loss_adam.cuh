// Loss and optimiser step of the training inner loop (sm_100a):
//   softmax_ce_kernel   CrossEntropyLoss(reduction='sum') forward + gradient in one pass
//                       (/root/reference/train.py:36,82: loss_fn(logits, labels[batch_nids]))
//   adam_kernel         torch.optim.Adam(lr, weight_decay) step, L2 decay folded into the gradient
//                       (/root/reference/train.py:34-35,85)
// Both are tiny next to the aggregation; they exist so that the whole step of SURVEY §8a row 9 runs
// in this library, and are bit-compatible with torch up to fp32 rounding of exp/log/sqrt.
#pragma once
#include "common.cuh"

namespace wsage {

// One warp per row (K classes, K <= 1024).  loss_partial[blockIdx.x] receives the block's loss sum; the
// host-side wrapper adds the partials in index order (deterministic).
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int64_t m, int k,
                  float* __restrict__ dlogits, int64_t ld_d, float* __restrict__ loss_partial) {
    __shared__ float warp_loss[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float my_loss = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < m; r += (int64_t)gridDim.x * 8) {
        const float* row = logits + r * ld;
        float mx = -INFINITY;
        for (int c = lane; c < k; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int c = lane; c < k; c += 32) sum += expf(row[c] - mx);
        sum = warp_sum(sum);
        const int64_t y = labels[r];
        const float lse = mx + logf(sum);
        // a label outside [0, k) (e.g. the -1 of a gene row) must not read out of bounds: it poisons the loss instead
        if (lane == 0) my_loss += (y >= 0 && y < k) ? lse - row[y] : NAN;
        if (dlogits) {
            const float inv = 1.f / sum;
            for (int c = lane; c < k; c += 32)
                dlogits[r * ld_d + c] = expf(row[c] - mx) * inv - (c == y ? 1.f : 0.f);
        }
    }
    if (lane == 0) warp_loss[warp] = my_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += warp_loss[w];
        loss_partial[blockIdx.x] = s;
    }
}

// p, g, m, v: n floats.  step_size = lr / (1 - beta1^step), 1 - beta and sqrt(1 - beta2^step) are formed on the
// host in fp64 (as torch does) and passed as fp32 scalars.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float step_size, float beta1, float omb1, float beta2, float omb2, float eps, float weight_decay, float bc2_sqrt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float grad = g[i];
        const float w = p[i];
        if (weight_decay != 0.f) grad = fmaf(weight_decay, w, grad);
        const float mi = fmaf(omb1, grad - m[i], m[i]);                 // torch: m.lerp_(grad, 1 - beta1)
        const float vi = fmaf(omb2 * grad, grad, beta2 * v[i]);         // torch: v.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = w - step_size * (mi / denom);
    }
}

}  // namespace wsage
