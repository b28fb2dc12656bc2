// Multi-tensor Adam: one launch updates every parameter of an optimizer (reference: torch.optim.Adam as built by
// models/base_model.py:39-48 -- lr/2 and 2*lr with betas (0, 0.999) under TTUR -- stepped twice per iteration,
// loss_collector.py:217-228).  The ~600 parameter tensors of G (98 M elements) are described once by a device-resident table
// (param / grad / exp_avg / exp_avg_sq pointers + element counts, cut into chunks of ADAM_CHUNK elements so the grid is even);
// the step counter lives on the device, so the update is capturable into the training step's CUDA graph.
//
// Arithmetic follows torch.optim.Adam (no amsgrad, no weight decay, maximize=False), fp32:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Roofline: HBM stream, 4 reads + 3 writes of 4 bytes per element (28 B/element: 2.7 GB per generator step).
#include "common.cuh"

#define ADAM_CHUNK 4096
#define ADAM_THREADS 256

__global__ void __launch_bounds__(ADAM_THREADS) k_adam(const fsv_adam_item* __restrict__ items, const int2* __restrict__ chunks,
                                                       float* __restrict__ step, float lr, float b1, float b2, float eps, int advance) {
    const int2 ch = chunks[blockIdx.x];                 // (item, chunk index inside the item)
    const fsv_adam_item it = items[ch.x];
    // every block computes the same bias corrections from the device-side step counter; block 0 advances it afterwards is a
    // race, so the counter is advanced by a separate 1-thread kernel launched first (k_adam_tick)
    const float t = *step;
    const float bc1 = 1.f - powf(b1, t);
    const float bc2s = sqrtf(1.f - powf(b2, t));
    const float step_size = lr / bc1;
    const long long base = (long long)ch.y * ADAM_CHUNK;
    const long long end = min(base + (long long)ADAM_CHUNK, it.numel);
    float* p = it.param; const float* g = it.grad; float* m = it.exp_avg; float* v = it.exp_avg_sq;
    (void)advance;
    if (((it.numel & 3) == 0) && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0)) {
        for (long long i = base + 4 * threadIdx.x; i < end; i += 4 * ADAM_THREADS) {
            float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ma[k] = b1 * ma[k] + (1.f - b1) * ga[k];
                va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];
                pa[k] -= step_size * (ma[k] / (sqrtf(va[k]) / bc2s + eps));
            }
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        }
    } else {
        for (long long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
            float gv = g[i];
            float mv = b1 * m[i] + (1.f - b1) * gv;
            float vv = b2 * v[i] + (1.f - b2) * gv * gv;
            m[i] = mv; v[i] = vv;
            p[i] -= step_size * (mv / (sqrtf(vv) / bc2s + eps));
        }
    }
}

__global__ void k_adam_tick(float* step) { *step += 1.f; }

extern "C" long long fsv_adam_chunks(long long numel) { return (numel + ADAM_CHUNK - 1) / ADAM_CHUNK; }

extern "C" int fsv_adam_step(const fsv_adam_item* items_dev, const int* chunks_dev, long long nchunks, float* step_dev, float lr,
                             float beta1, float beta2, float eps, void* stream) {
    FSV_REQUIRE(items_dev && chunks_dev && step_dev && nchunks > 0, "adam_step: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    k_adam_tick<<<1, 1, 0, st>>>(step_dev);             // t <- t + 1 first, as torch does (state['step'] += 1 before the update)
    FSV_CHECK_LAUNCH("adam_tick");
    k_adam<<<(unsigned)nchunks, ADAM_THREADS, 0, st>>>(items_dev, reinterpret_cast<const int2*>(chunks_dev), step_dev, lr, beta1, beta2, eps, 1);
    FSV_CHECK_LAUNCH("adam_step");
    return FSV_OK;
}
