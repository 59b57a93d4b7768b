// abx_qnet.cu -- batched Q-network forward of the DDQN execution agent on the 5th-generation tensor cores (sm_100a).
//
// Reference: util/model/QNets.py:7-27,55-60 (EvalModel / TargetModel = NNModel_1: Dense 32-64-128-128-64-32 with ReLU, linear
// output of n_actions = 24; Dropout is the identity at inference) evaluated by Keras `predict` on a batch of ONE state per decision
// tick (agent/execution/qlearning/ddqlearning_execution_agent.py:339-365); here one launch evaluates every environment of the batch
// and applies the agent's action rule (np.argmax, or the epsilon branch of choose_action :349-357 with Philox instead of np.random).
//
// One persistent CTA per SM walks 128-row tiles of the batch:
//   * the whole network (all layers, fp32 split into bf16 hi + bf16 lo parts) is brought into shared memory once per CTA by
//     per-layer 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), so layer 0 starts while later layers still stream;
//   * each layer is D[128 x N] = A[128 x K] . W[N x K]^T on tcgen05.mma (kind::f16, bf16 operands, fp32 accumulator in TMEM),
//     issued by one thread; three MMAs per 16-wide k step (hi.hi + hi.lo + lo.hi) give fp32-class accuracy (~1e-5 relative) --
//     the reference runs the network in fp32, and a flipped argmax is a different order in the book;
//   * the epilogue (4 warps = 128 TMEM lanes = 128 rows) reads the accumulator with tcgen05.ld, adds the bias, applies ReLU,
//     splits into hi/lo bf16 and writes the next layer's A operand straight back to shared memory in the canonical K-major
//     core-matrix layout (no swizzle: 8 rows x 16 B core matrices, 16-byte stores, conflict free); activations never touch HBM;
//   * the last layer's epilogue does the argmax / epsilon rule and writes one int32 action (and optionally the fp32 Q row).
// HBM traffic per row: 16 B of state in, 4 B of action out (+ 96 B when Q is requested); the weights (~150 KB) are read once per CTA
// from L2.  The kernel is latency bound by construction (7 dependent layers per tile); it exists so that acting costs a few
// microseconds per tick instead of 8 192 Keras calls.
//
// Networks that do not fit (util/model/QNets.py:30-52 NNModel_2: ... 128-256-128 ..., 347 KB of hi + lo weights) run through abx_qnet_forward_streamed_kernel: the same
// tile loop, MMA sequence and epilogue, but every layer is cut into blocks of at most 128 x 128 weights (64 KB hi + lo) that are streamed from L2 into ONE shared-memory
// buffer per tile -- N blocks land in their own 128 TMEM columns, K blocks accumulate -- while the A operand holds up to 256 features per row (2 x 64 KB).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "../../include/abides_b200.h"

namespace {

constexpr int QN_MAX_LAYERS = 8, QN_TILE_M = 128, QN_MAX_DIM = 128, QN_THREADS = 128, QN_TMEM_COLS = 128;
constexpr int QN_S_MAX_DIM = 256, QN_S_TMEM_COLS = 256, QN_BLK = 128;                     // streamed mode: layer widths up to 256, weight blocks of at most QN_BLK x QN_BLK

struct QnetDev {
  int32_t n_layers, n_in, n_out, pad0;
  int32_t kpad[QN_MAX_LAYERS], npad[QN_MAX_LAYERS];
  uint32_t w_off[QN_MAX_LAYERS];       // byte offset of layer l's image (hi block then lo block) inside the weight image
  uint32_t b_off[QN_MAX_LAYERS];       // float offset of layer l's bias inside the bias array
  uint32_t w_bytes, n_bias;
  int32_t kblk[QN_MAX_LAYERS], nblk[QN_MAX_LAYERS];   // weight block dims of layer l: the whole layer (resident mode) or min(dim, QN_BLK) (streamed mode); block (nb, kb) starts at w_off + (nb * (kpad / kblk) + kb) * 4 * nblk * kblk bytes: hi image then lo image, K-major core matrices
};

// Philox4x32-10 (same generator as the simulator's streams, abx_core.cuh), for the epsilon branch of choose_action
__device__ __forceinline__ uint4 qn_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
    uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0, n1 = uint32_t(p1), n2 = uint32_t(p0 >> 32) ^ c3 ^ k1, n3 = uint32_t(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading (K-chunk) and
// stride (8-row group) byte offsets in 16-byte units, descriptor version 1 (Blackwell), layout type 0 (SWIZZLE_NONE).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, dense, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(QN_TILE_M >> 4) << 24); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                 "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) { hi = __float2bfloat16_rn(x); lo = __float2bfloat16_rn(x - __bfloat162float(hi)); }
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) { return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16); }

// Shared memory: [weight image | A hi (128 x 128 bf16) | A lo | biases | mbarriers | TMEM base]
__global__ void __launch_bounds__(QN_THREADS, 1)
abx_qnet_forward_kernel(QnetDev net, const uint8_t *__restrict__ wimg, const float *__restrict__ bias, const double *__restrict__ x, int x_stride, int x_offset, int n_rows,
                        float *__restrict__ q_out, int32_t *__restrict__ action_out, double greedy_prob, uint64_t seed, uint64_t counter) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t *s_w = smem;
  uint8_t *s_ahi = smem + net.w_bytes, *s_alo = s_ahi + QN_TILE_M * QN_MAX_DIM * 2;
  float *s_bias = reinterpret_cast<float *>(s_alo + QN_TILE_M * QN_MAX_DIM * 2);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_bias + ((net.n_bias + 3) & ~3u));       // [0..L) weights of layer l landed, [L] MMAs of a layer retired
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + QN_MAX_LAYERS + 1);
  const uint32_t bar0 = smem_u32(s_bar), mma_bar = bar0 + 8u * net.n_layers;

  if (tid == 0) {
    for (int l = 0; l <= net.n_layers; l++) mbar_init(bar0 + 8u * l, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                                                                          // one warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)QN_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = tid; i < net.n_bias; i += QN_THREADS) s_bias[i] = bias[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  if (tid == 0) {                                                                           // stream the network: one bulk copy per layer
    for (int l = 0; l < net.n_layers; l++) {
      uint32_t bytes = 2u * (uint32_t)net.npad[l] * (uint32_t)net.kpad[l] * 2u;
      mbar_expect_tx(bar0 + 8u * l, bytes);
      bulk_g2s(smem_u32(s_w + net.w_off[l]), wimg + net.w_off[l], bytes, bar0 + 8u * l);
    }
  }
  uint32_t phase = 0; bool first = true;
  const int n_tiles = (n_rows + QN_TILE_M - 1) / QN_TILE_M;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row = tile * QN_TILE_M + tid;
    {                                                                                       // layer-0 operand: the state, zero padded to kpad[0]
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; k++) v[k] = (row < n_rows && k < net.n_in) ? (float)x[(size_t)row * x_stride + x_offset + k] : 0.0f;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        uint32_t h[4], lo4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { __nv_bfloat16 h0, l0, h1, l1; split_bf16(v[8 * c + 2 * j], h0, l0); split_bf16(v[8 * c + 2 * j + 1], h1, l1); h[j] = pack2(h0, h1); lo4[j] = pack2(l0, l1); }
        *reinterpret_cast<uint4 *>(s_ahi + c * (QN_TILE_M * 16) + tid * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(s_alo + c * (QN_TILE_M * 16) + tid * 16) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
      }
    }
#pragma unroll 1
    for (int l = 0; l < net.n_layers; l++) {
      const int K = net.kpad[l], N = net.npad[l];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                          // generic-proxy stores of A -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        if (first) mbar_wait(bar0 + 8u * l, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = umma_idesc_bf16(N);
        const uint32_t a_lbo = QN_TILE_M * 16, b_lbo = (uint32_t)N * 16, sbo = 128;
        const uint32_t a_hi = smem_u32(s_ahi), a_lo = smem_u32(s_alo), b_hi = smem_u32(s_w + net.w_off[l]), b_lo = b_hi + (uint32_t)N * (uint32_t)K * 2u;
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ks++) {
          const uint64_t dah = umma_desc(a_hi + ks * 2 * a_lbo, a_lbo, sbo), dal = umma_desc(a_lo + ks * 2 * a_lbo, a_lbo, sbo);
          const uint64_t dbh = umma_desc(b_hi + ks * 2 * b_lbo, b_lbo, sbo), dbl = umma_desc(b_lo + ks * 2 * b_lbo, b_lbo, sbo);
          umma_bf16(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);                                // small terms first
          umma_bf16(tmem, dah, dbl, idesc, 1u);
          umma_bf16(tmem, dah, dbh, idesc, 1u);
        }
        umma_commit(mma_bar);                                                               // arrives when every MMA above has retired
      }
      mbar_wait(mma_bar, phase); phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);                          // warp w reads TMEM lanes 32w .. 32w+31 (row = lane)
      const float *b = s_bias + net.b_off[l];
      if (l + 1 < net.n_layers) {
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16]; tmem_ld16(taddr + (uint32_t)c0, v);
          uint32_t h[8], lo8[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            float f0 = fmaxf(__uint_as_float(v[2 * j]) + b[c0 + 2 * j], 0.0f), f1 = fmaxf(__uint_as_float(v[2 * j + 1]) + b[c0 + 2 * j + 1], 0.0f);
            __nv_bfloat16 h0, l0, h1, l1; split_bf16(f0, h0, l0); split_bf16(f1, h1, l1); h[j] = pack2(h0, h1); lo8[j] = pack2(l0, l1);
          }
          const int ch = c0 >> 3;                                                           // next layer's K chunk (8 features = 16 bytes per row)
          *reinterpret_cast<uint4 *>(s_ahi + ch * (QN_TILE_M * 16) + tid * 16) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4 *>(s_ahi + (ch + 1) * (QN_TILE_M * 16) + tid * 16) = make_uint4(h[4], h[5], h[6], h[7]);
          *reinterpret_cast<uint4 *>(s_alo + ch * (QN_TILE_M * 16) + tid * 16) = make_uint4(lo8[0], lo8[1], lo8[2], lo8[3]);
          *reinterpret_cast<uint4 *>(s_alo + (ch + 1) * (QN_TILE_M * 16) + tid * 16) = make_uint4(lo8[4], lo8[5], lo8[6], lo8[7]);
        }
      } else {                                                                              // output layer: Q row, argmax / epsilon rule
        float best = -INFINITY; int best_i = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16]; tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 16; j++) {
            int col = c0 + j;
            if (col < net.n_out) {
              float q = __uint_as_float(v[j]) + b[col];
              if (q_out && row < n_rows) q_out[(size_t)row * net.n_out + col] = q;
              if (q > best) { best = q; best_i = col; }                                     // first maximum, like np.argmax
            }
          }
        }
        if (action_out && row < n_rows) {
          int a = best_i;
          if (greedy_prob < 1.0) {                                                          // choose_action :349-357: greedy with probability epsilon, else uniform
            uint4 r = qn_philox((uint32_t)row, (uint32_t)counter, (uint32_t)(counter >> 32), 0x514e4554u, (uint32_t)seed, (uint32_t)(seed >> 32));
            double u = ((r.x >> 5) * 67108864.0 + (r.y >> 6)) / 9007199254740992.0;
            if (!(u < greedy_prob)) a = (int)(((uint64_t)r.z * (uint64_t)net.n_out) >> 32);
          }
          action_out[row] = a;
        }
      }
    }
    first = false;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && first) for (int l = 0; l < net.n_layers; l++) mbar_wait(bar0 + 8u * l, 0);   // a CTA without tiles still has to let its bulk copies land
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)QN_TMEM_COLS) : "memory");
}

// Streamed mode (see the header comment).  Shared memory: [one weight block, 64 KB | A hi (128 x 256 bf16) | A lo | biases | 2 mbarriers | TMEM base]
__global__ void __launch_bounds__(QN_THREADS, 1)
abx_qnet_forward_streamed_kernel(QnetDev net, const uint8_t *__restrict__ wimg, const float *__restrict__ bias, const double *__restrict__ x, int x_stride, int x_offset, int n_rows,
                                 float *__restrict__ q_out, int32_t *__restrict__ action_out, double greedy_prob, uint64_t seed, uint64_t counter) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t *s_w = smem;
  uint8_t *s_ahi = smem + 4 * QN_BLK * QN_BLK, *s_alo = s_ahi + QN_TILE_M * QN_S_MAX_DIM * 2;
  float *s_bias = reinterpret_cast<float *>(s_alo + QN_TILE_M * QN_S_MAX_DIM * 2);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_bias + ((net.n_bias + 3) & ~3u));       // [0] a weight block landed, [1] the MMAs of a block retired
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2);
  const uint32_t w_bar = smem_u32(s_bar), mma_bar = w_bar + 8u;
  if (tid == 0) { mbar_init(w_bar, 1); mbar_init(mma_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)QN_S_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = tid; i < net.n_bias; i += QN_THREADS) s_bias[i] = bias[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  uint32_t wphase = 0, mphase = 0;
  const int n_tiles = (n_rows + QN_TILE_M - 1) / QN_TILE_M;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row = tile * QN_TILE_M + tid;
    {                                                                                       // layer-0 operand: the state, zero padded to kpad[0] (= 16)
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; k++) v[k] = (row < n_rows && k < net.n_in) ? (float)x[(size_t)row * x_stride + x_offset + k] : 0.0f;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        uint32_t h[4], lo4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { __nv_bfloat16 h0, l0, h1, l1; split_bf16(v[8 * c + 2 * j], h0, l0); split_bf16(v[8 * c + 2 * j + 1], h1, l1); h[j] = pack2(h0, h1); lo4[j] = pack2(l0, l1); }
        *reinterpret_cast<uint4 *>(s_ahi + c * (QN_TILE_M * 16) + tid * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(s_alo + c * (QN_TILE_M * 16) + tid * 16) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
      }
    }
#pragma unroll 1
    for (int l = 0; l < net.n_layers; l++) {
      const int K = net.kpad[l], N = net.npad[l], Kb = net.kblk[l], Nb = net.nblk[l], nkb = K / Kb, nnb = N / Nb;
      const uint32_t blk_bytes = 4u * (uint32_t)Nb * (uint32_t)Kb;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                          // generic-proxy stores of A -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
#pragma unroll 1
      for (int nb = 0; nb < nnb; nb++) {
#pragma unroll 1
        for (int kb = 0; kb < nkb; kb++) {
          if (tid == 0) {                                                                   // the previous block's MMAs have retired (mma_bar below): the buffer is free
            mbar_expect_tx(w_bar, blk_bytes);
            bulk_g2s(smem_u32(s_w), wimg + net.w_off[l] + (size_t)(nb * nkb + kb) * blk_bytes, blk_bytes, w_bar);
            mbar_wait(w_bar, wphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = umma_idesc_bf16(Nb);
            const uint32_t a_lbo = QN_TILE_M * 16, b_lbo = (uint32_t)Nb * 16, sbo = 128;
            const uint32_t a_hi = smem_u32(s_ahi) + (uint32_t)(kb * Kb / 8) * a_lbo, a_lo = smem_u32(s_alo) + (uint32_t)(kb * Kb / 8) * a_lbo;
            const uint32_t b_hi = smem_u32(s_w), b_lo = b_hi + (uint32_t)Nb * (uint32_t)Kb * 2u, d = tmem + (uint32_t)(nb * Nb);
#pragma unroll 1
            for (int ks = 0; ks < Kb / 16; ks++) {
              const uint64_t dah = umma_desc(a_hi + ks * 2 * a_lbo, a_lbo, sbo), dal = umma_desc(a_lo + ks * 2 * a_lbo, a_lbo, sbo);
              const uint64_t dbh = umma_desc(b_hi + ks * 2 * b_lbo, b_lbo, sbo), dbl = umma_desc(b_lo + ks * 2 * b_lbo, b_lbo, sbo);
              umma_bf16(d, dal, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);                   // small terms first
              umma_bf16(d, dah, dbl, idesc, 1u);
              umma_bf16(d, dah, dbh, idesc, 1u);
            }
            umma_commit(mma_bar);
          }
          wphase ^= 1u;
          mbar_wait(mma_bar, mphase); mphase ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
      }
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      const float *b = s_bias + net.b_off[l];
      if (l + 1 < net.n_layers) {
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16]; tmem_ld16(taddr + (uint32_t)c0, v);
          uint32_t h[8], lo8[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            float f0 = fmaxf(__uint_as_float(v[2 * j]) + b[c0 + 2 * j], 0.0f), f1 = fmaxf(__uint_as_float(v[2 * j + 1]) + b[c0 + 2 * j + 1], 0.0f);
            __nv_bfloat16 h0, l0, h1, l1; split_bf16(f0, h0, l0); split_bf16(f1, h1, l1); h[j] = pack2(h0, h1); lo8[j] = pack2(l0, l1);
          }
          const int ch = c0 >> 3;
          *reinterpret_cast<uint4 *>(s_ahi + ch * (QN_TILE_M * 16) + tid * 16) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4 *>(s_ahi + (ch + 1) * (QN_TILE_M * 16) + tid * 16) = make_uint4(h[4], h[5], h[6], h[7]);
          *reinterpret_cast<uint4 *>(s_alo + ch * (QN_TILE_M * 16) + tid * 16) = make_uint4(lo8[0], lo8[1], lo8[2], lo8[3]);
          *reinterpret_cast<uint4 *>(s_alo + (ch + 1) * (QN_TILE_M * 16) + tid * 16) = make_uint4(lo8[4], lo8[5], lo8[6], lo8[7]);
        }
      } else {
        float best = -INFINITY; int best_i = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16]; tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 16; j++) {
            int col = c0 + j;
            if (col < net.n_out) {
              float q = __uint_as_float(v[j]) + b[col];
              if (q_out && row < n_rows) q_out[(size_t)row * net.n_out + col] = q;
              if (q > best) { best = q; best_i = col; }
            }
          }
        }
        if (action_out && row < n_rows) {
          int a = best_i;
          if (greedy_prob < 1.0) {
            uint4 r = qn_philox((uint32_t)row, (uint32_t)counter, (uint32_t)(counter >> 32), 0x514e4554u, (uint32_t)seed, (uint32_t)(seed >> 32));
            double u = ((r.x >> 5) * 67108864.0 + (r.y >> 6)) / 9007199254740992.0;
            if (!(u < greedy_prob)) a = (int)(((uint64_t)r.z * (uint64_t)net.n_out) >> 32);
          }
          action_out[row] = a;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)QN_S_TMEM_COLS) : "memory");
}

// Device-side version of pack_params: fp32 parameters (per layer W[out][in] then b[out]) -> bf16 hi/lo image in the MMA operand layout + bias
// array, so that the learner can push new weights without a host round trip.  One thread per padded weight / bias element.
__global__ void abx_qnet_pack_kernel(QnetDev net, int d0, int d1, int d2, int d3, int d4, int d5, int d6, int d7, int d8, const float *__restrict__ params,
                                     uint8_t *__restrict__ wimg, float *__restrict__ bias) {
  const int dims[9] = {d0, d1, d2, d3, d4, d5, d6, d7, d8};
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0, poff = 0;
  for (int l = 0; l < net.n_layers; l++) {
    uint32_t K = net.kpad[l], N = net.npad[l], nw = N * K, in = dims[l], out = dims[l + 1];
    if (t < acc + nw) {                                                   // weight element (o, i) of the padded layer
      uint32_t e = t - acc, o = e / K, i = e % K;
      float w = (o < out && i < in) ? params[poff + o * in + i] : 0.0f;
      __nv_bfloat16 hi, lo; split_bf16(w, hi, lo);
      uint32_t Kb = net.kblk[l], Nb = net.nblk[l], nb = o / Nb, kb = i / Kb, oo = o % Nb, ii = i % Kb;   // block (nb, kb); resident mode: one block = the layer
      uint32_t dst = (ii >> 3) * (Nb * 8) + oo * 8 + (ii & 7);
      __nv_bfloat16 *base = reinterpret_cast<__nv_bfloat16 *>(wimg + net.w_off[l] + (size_t)(nb * (K / Kb) + kb) * 4u * Nb * Kb);
      base[dst] = hi; base[(size_t)Nb * Kb + dst] = lo;
      return;
    }
    acc += nw;
    if (t < acc + N) { uint32_t o = t - acc; bias[net.b_off[l] + o] = o < out ? params[poff + out * in + o] : 0.0f; return; }
    acc += N; poff += out * in + out;
  }
}

thread_local char g_err[512] = "";

}  // namespace

struct abx_qnet {
  QnetDev net; bool streamed; int device, n_sms; size_t smem_bytes; uint8_t *d_wimg; float *d_bias; std::vector<int32_t> dims; int64_t launches;
  std::vector<uint8_t> h_wimg; std::vector<float> h_bias;
};

#define QCU(call)                                                                                          \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); return ABX_ERR_CUDA; } } while (0)

static inline int pad16(int v) { return (v + 15) & ~15; }
static inline uint16_t f2bf(float f) {                                                        // round to nearest even, like __float2bfloat16_rn
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

// params: for every layer W[out][in] (row major, i.e. the transpose of a Keras Dense kernel) followed by b[out]
static void pack_params(abx_qnet *q, const float *params) {
  const QnetDev &n = q->net; const float *p = params;
  std::fill(q->h_wimg.begin(), q->h_wimg.end(), 0); std::fill(q->h_bias.begin(), q->h_bias.end(), 0.0f);
  for (int l = 0; l < n.n_layers; l++) {
    int in = q->dims[l], out = q->dims[l + 1], K = n.kpad[l], N = n.npad[l];
    const int Kb = n.kblk[l], Nb = n.nblk[l]; (void)N;
    for (int o = 0; o < out; o++) for (int i = 0; i < in; i++) {
      float w = p[(size_t)o * in + i]; uint16_t h = f2bf(w), lw = f2bf(w - bf2f(h));
      int nb = o / Nb, kb = i / Kb, oo = o % Nb, ii = i % Kb;                                 // block (nb, kb): the whole layer in resident mode
      uint16_t *hi = reinterpret_cast<uint16_t *>(q->h_wimg.data() + n.w_off[l] + (size_t)(nb * (K / Kb) + kb) * 4u * Nb * Kb), *lo = hi + (size_t)Nb * Kb;
      size_t e = (size_t)(ii >> 3) * ((size_t)Nb * 8) + (size_t)oo * 8 + (ii & 7);            // K chunk (8 elements = 16 B per row), then row, then element
      hi[e] = h; lo[e] = lw;
    }
    p += (size_t)out * in;
    for (int o = 0; o < out; o++) q->h_bias[n.b_off[l] + o] = p[o];
    p += out;
  }
}

extern "C" {

const char *abx_qnet_last_error(void) { return g_err; }

int32_t abx_qnet_param_count(const int32_t *dims, int32_t n_layers) {
  if (!dims || n_layers < 1 || n_layers > QN_MAX_LAYERS) return -1;
  int64_t t = 0; for (int l = 0; l < n_layers; l++) t += (int64_t)dims[l] * dims[l + 1] + dims[l + 1];
  return t > 0x7fffffff ? -1 : (int32_t)t;
}

int32_t abx_qnet_create(const int32_t *dims, int32_t n_layers, const float *params, int32_t device, abx_qnet **out) {
  if (!out || !dims || !params || n_layers < 1 || n_layers > QN_MAX_LAYERS) return ABX_ERR_ARG;
  if (dims[0] < 1 || dims[0] > 16) return ABX_ERR_ARG;                                       // the state is staged as one 16-wide k step
  bool wide = false;
  for (int l = 1; l <= n_layers; l++) { if (dims[l] < 1 || dims[l] > QN_S_MAX_DIM) return ABX_ERR_ARG; if (dims[l] > QN_MAX_DIM) wide = true; }
  int ndev = 0; QCU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { snprintf(g_err, sizeof(g_err), "device %d not present (%d visible)", device, ndev); return ABX_ERR_CUDA; }
  QCU(cudaSetDevice(device));
  abx_qnet *q = new (std::nothrow) abx_qnet(); if (!q) return ABX_ERR_ARG;
  q->device = device; q->launches = 0; q->d_wimg = nullptr; q->d_bias = nullptr; q->dims.assign(dims, dims + n_layers + 1);
  QnetDev &n = q->net; memset(&n, 0, sizeof(n));
  n.n_layers = n_layers; n.n_in = dims[0]; n.n_out = dims[n_layers];
  uint32_t woff = 0, boff = 0;
  for (int l = 0; l < n_layers; l++) {
    n.kpad[l] = dims[l] > QN_MAX_DIM ? QN_S_MAX_DIM : pad16(dims[l]); n.npad[l] = dims[l + 1] > QN_MAX_DIM ? QN_S_MAX_DIM : pad16(dims[l + 1]); if (l == n_layers - 1 && n.npad[l] < 32) n.npad[l] = 32;
    n.w_off[l] = woff; n.b_off[l] = boff; woff += 2u * n.npad[l] * n.kpad[l] * 2u; boff += n.npad[l];
  }
  n.w_bytes = woff; n.n_bias = boff;
  cudaDeviceProp prop; QCU(cudaGetDeviceProperties(&prop, device)); q->n_sms = prop.multiProcessorCount;
  q->smem_bytes = (size_t)woff + 2 * QN_TILE_M * QN_MAX_DIM * 2 + (size_t)((boff + 3) & ~3u) * 4 + (QN_MAX_LAYERS + 1) * 8 + 16;
  { const char *f = getenv("ABX_QNET_FORCE_STREAMED"); if (f && f[0] == '1') wide = true; }   // test hook: run a network that would fit through the streamed kernel
  q->streamed = wide || q->smem_bytes > (size_t)prop.sharedMemPerBlockOptin;                 // the resident kernel needs every layer <= 128 wide and the whole image on chip
  for (int l = 0; l < n_layers; l++) { n.kblk[l] = (q->streamed && n.kpad[l] > QN_BLK) ? QN_BLK : n.kpad[l]; n.nblk[l] = (q->streamed && n.npad[l] > QN_BLK) ? QN_BLK : n.npad[l]; }
  if (q->streamed) q->smem_bytes = (size_t)4 * QN_BLK * QN_BLK + 2 * QN_TILE_M * QN_S_MAX_DIM * 2 + (size_t)((boff + 3) & ~3u) * 4 + 2 * 8 + 16;
  if (q->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) { snprintf(g_err, sizeof(g_err), "network needs %zu B of shared memory (limit %zu)", q->smem_bytes, (size_t)prop.sharedMemPerBlockOptin); delete q; return ABX_ERR_ARG; }
  q->h_wimg.resize(woff); q->h_bias.resize(boff);
  if (cudaMalloc((void **)&q->d_wimg, woff) != cudaSuccess || cudaMalloc((void **)&q->d_bias, sizeof(float) * boff) != cudaSuccess) { cudaFree(q->d_wimg); delete q; snprintf(g_err, sizeof(g_err), "cudaMalloc failed"); return ABX_ERR_CUDA; }
  // from here on a failing CUDA call must not leak the handle or its device buffers
#define QCUH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s: %s", #call, cudaGetErrorString(e_)); cudaFree(q->d_wimg); cudaFree(q->d_bias); delete q; return ABX_ERR_CUDA; } } while (0)
  if (q->streamed) QCUH(cudaFuncSetAttribute(abx_qnet_forward_streamed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q->smem_bytes));
  else QCUH(cudaFuncSetAttribute(abx_qnet_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q->smem_bytes));
  pack_params(q, params);
  QCUH(cudaMemcpy(q->d_wimg, q->h_wimg.data(), woff, cudaMemcpyHostToDevice));
  QCUH(cudaMemcpy(q->d_bias, q->h_bias.data(), sizeof(float) * boff, cudaMemcpyHostToDevice));
#undef QCUH
  *out = q; return ABX_OK;
}

int32_t abx_qnet_set_params(abx_qnet *q, const float *params, void *stream) {
  if (!q || !params) return ABX_ERR_ARG;
  QCU(cudaSetDevice(q->device)); cudaStream_t st = (cudaStream_t)stream;
  QCU(cudaStreamSynchronize(st));                                                            // the staging image below is reused: earlier copies must have left it
  pack_params(q, params);
  QCU(cudaMemcpyAsync(q->d_wimg, q->h_wimg.data(), q->net.w_bytes, cudaMemcpyHostToDevice, st));
  QCU(cudaMemcpyAsync(q->d_bias, q->h_bias.data(), sizeof(float) * q->net.n_bias, cudaMemcpyHostToDevice, st));
  QCU(cudaStreamSynchronize(st));
  return ABX_OK;
}

int32_t abx_qnet_set_params_device(abx_qnet *q, const float *params_dev, void *stream) {
  if (!q || !params_dev) return ABX_ERR_ARG;
  QCU(cudaSetDevice(q->device));
  int d[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; for (size_t i = 0; i < q->dims.size() && i < 9; i++) d[i] = q->dims[i];
  uint32_t total = 0; for (int l = 0; l < q->net.n_layers; l++) total += (uint32_t)q->net.npad[l] * q->net.kpad[l] + q->net.npad[l];
  abx_qnet_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(q->net, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], params_dev, q->d_wimg, q->d_bias);
  q->launches += 1;
  QCU(cudaGetLastError());
  return ABX_OK;
}

int32_t abx_qnet_destroy(abx_qnet *q) { if (!q) return ABX_OK; cudaSetDevice(q->device); cudaFree(q->d_wimg); cudaFree(q->d_bias); delete q; return ABX_OK; }
int64_t abx_qnet_launch_count(const abx_qnet *q) { return q ? q->launches : 0; }

int32_t abx_qnet_forward(abx_qnet *q, const double *x_dev, int32_t x_stride, int32_t x_offset, int32_t n, float *q_out_dev, int32_t *action_out_dev,
                         double greedy_prob, uint64_t seed, uint64_t counter, void *stream) {
  if (!q || !x_dev || n < 1 || x_stride < 1 || x_offset < 0 || x_offset + q->net.n_in > x_stride || (!q_out_dev && !action_out_dev)) return ABX_ERR_ARG;
  QCU(cudaSetDevice(q->device));
  int tiles = (n + QN_TILE_M - 1) / QN_TILE_M, grid = tiles < q->n_sms ? tiles : q->n_sms;
  if (q->streamed) abx_qnet_forward_streamed_kernel<<<grid, QN_THREADS, q->smem_bytes, (cudaStream_t)stream>>>(q->net, q->d_wimg, q->d_bias, x_dev, x_stride, x_offset, n, q_out_dev, action_out_dev, greedy_prob, seed, counter);
  else abx_qnet_forward_kernel<<<grid, QN_THREADS, q->smem_bytes, (cudaStream_t)stream>>>(q->net, q->d_wimg, q->d_bias, x_dev, x_stride, x_offset, n, q_out_dev, action_out_dev, greedy_prob, seed, counter);
  q->launches += 1;
  QCU(cudaGetLastError());
  return ABX_OK;
}

}  // extern "C"
