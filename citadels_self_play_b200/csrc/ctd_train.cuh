// ctd_train.cuh -- the value network's training step on the device (algorithms/train.py:13-86 `train_node_value_only`):
// ValueOnlyNN(418, 512) in train mode (BatchNorm1d with batch statistics, Dropout(0.2)), KLDivLoss(batchmean) between
// log(square_and_normalize(outputs) + 1e-10) and square_and_normalize(labels) (train_utils.py:143-152), Adam.
//
// The three dense layers' products -- forward X.W^T, backward-data dZ.W, backward-weight dZ^T.X -- all run on the tensor-core
// kernel of the inference path (ctd_k_linear_tc: C = A.B^T with both operands row-major over K, 3xTF32 split precision, fp32
// accumulation in TMEM); operands that are needed the other way round are transposed by a small kernel first.  Everything
// around them (batch gather, BatchNorm forward / backward over the batch, ReLU, dropout, fc4, the loss and its gradient, bias
// gradients, Adam) is plain fp32 CUDA-core code; the loss itself is formed in fp64 because the reference's labels are float64.
//
// Chance is counter based like everywhere else in the engine: the dropout mask of element e of layer L at optimiser step s is
// bit-defined by Philox4x32-10 (key = seed, counter = (e >> 2, L | s << 8, 0xD0, 0), word e & 3): keep iff word < 0.8 * 2^32.
// The batch order of an epoch is a permutation handed in by the host (train.py draws it from the same generator), so a run is a
// pure function of (initial weights, data, seed) and can be replayed against the reference with its DataLoader / dropout routed
// through the same draws (tests/golden/gen_train_fixture.py).
#pragma once
#include <stdint.h>

#define CTD_TR_IN 512          /* 418 features padded to a multiple of 128: the padding columns of fc1 stay zero */
#define CTD_TR_H1 512
#define CTD_TR_H2 256
#define CTD_TR_H3 128
#define CTD_TR_KEEP_U32 3435973836u   /* floor(0.8 * 2^32): Dropout(0.2) keeps an element iff its Philox word is below this */
#define CTD_TR_BN_EPS 1e-5f
#define CTD_TR_BN_MOMENTUM 0.1f

// parameters and optimiser state, one flat fp32 array each (same offsets): fc weights [out][in] like nn.Linear
struct CtdTrainLayout {
  // offsets in floats
  static constexpr size_t w1 = 0, b1 = w1 + (size_t)CTD_TR_H1 * CTD_TR_IN, g1 = b1 + CTD_TR_H1, be1 = g1 + CTD_TR_H1,
                          w2 = be1 + CTD_TR_H1, b2 = w2 + (size_t)CTD_TR_H2 * CTD_TR_H1, g2 = b2 + CTD_TR_H2, be2 = g2 + CTD_TR_H2,
                          w3 = be2 + CTD_TR_H2, b3 = w3 + (size_t)CTD_TR_H3 * CTD_TR_H2, w4 = b3 + CTD_TR_H3,
                          b4 = w4 + 6 * CTD_TR_H3, total = ((b4 + 6 + 127) / 128) * 128;
};

__device__ __forceinline__ uint32_t ctd_tr_philox_word(uint64_t seed, uint32_t step, uint32_t layer, uint32_t e) {
  uint32_t c0 = e >> 2, c1 = layer | (step << 8), c2 = 0xD0u, c3 = 0u, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint32_t w = e & 3u;
  return w == 0 ? c0 : (w == 1 ? c1 : (w == 2 ? c2 : c3));
}

// rows perm[first .. first+B) of the data set -> X [Bp][512] (features 418.. and rows B.. zero), labels -> T [Bp][6] (fp64)
__global__ void ctd_k_tr_gather(const float* __restrict__ feats, const double* __restrict__ vals, const uint32_t* __restrict__ perm,
                                uint32_t first, uint32_t B, uint32_t Bp, float* __restrict__ X, double* __restrict__ T) {
  const uint32_t r = blockIdx.x;
  const uint32_t src = r < B ? (perm ? perm[first + r] : first + r) : 0;
  for (int c = threadIdx.x; c < CTD_TR_IN; c += blockDim.x) X[(size_t)r * CTD_TR_IN + c] = (r < B && c < 418) ? feats[(size_t)src * 418 + c] : 0.f;
  if (threadIdx.x < 6) T[(size_t)r * 6 + threadIdx.x] = r < B ? vals[(size_t)src * 6 + threadIdx.x] : 0.0;
}

// out [C][R] = in [R][C]^T  (R rows of ld_in floats); rows of `out` are padded with zeros up to ld_out >= R
__global__ void ctd_k_tr_transpose(const float* __restrict__ in, int R, int C, int ld_in, float* __restrict__ out, int ld_out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < ld_out) out[(size_t)c * ld_out + r] = tile[threadIdx.x][i];
  }
}

// BatchNorm1d (train mode) + ReLU + Dropout over the batch, one block of 32 x 8 threads per 32 columns:
//   mean / biased variance over the B real rows -> xhat = (z - mean) * invstd, a = gamma * xhat + beta, h = relu(a) * mask / keep
// Saves xhat (in place of z), invstd, and h; updates the running statistics (momentum 0.1, unbiased variance).  train == 0: eval
// mode (running statistics, no dropout).
__global__ void __launch_bounds__(256) ctd_k_tr_bn_fwd(float* __restrict__ Z, int N, uint32_t B, uint32_t Bp, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                                       float* __restrict__ invstd_out, float* __restrict__ H, int train, uint64_t seed,
                                                       uint32_t step, uint32_t layer) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y, col = blockIdx.x * 32 + tx;
  float mean, invstd;
  if (train) {
    float s = 0.f;
    for (uint32_t r = ty; r < B; r += 8) s += Z[(size_t)r * N + col];
    red[ty][tx] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    mean = s / (float)B;
    __syncthreads();
    float q = 0.f;
    for (uint32_t r = ty; r < B; r += 8) { const float d = Z[(size_t)r * N + col] - mean; q += d * d; }
    red[ty][tx] = q;
    __syncthreads();
    q = 0.f;
    for (int i = 0; i < 8; ++i) q += red[i][tx];
    const float var = q / (float)B;
    invstd = rsqrtf(var + CTD_TR_BN_EPS);
    if (ty == 0) {
      invstd_out[col] = invstd;
      rmean[col] = (1.f - CTD_TR_BN_MOMENTUM) * rmean[col] + CTD_TR_BN_MOMENTUM * mean;
      rvar[col] = (1.f - CTD_TR_BN_MOMENTUM) * rvar[col] + CTD_TR_BN_MOMENTUM * (B > 1 ? q / (float)(B - 1) : var);
    }
  } else {
    mean = rmean[col];
    invstd = rsqrtf(rvar[col] + CTD_TR_BN_EPS);
  }
  const float g = gamma[col], be = beta[col];
  for (uint32_t r = ty; r < Bp; r += 8) {
    const size_t i = (size_t)r * N + col;
    if (r >= B) { Z[i] = 0.f; H[i] = 0.f; continue; }
    const float xh = (Z[i] - mean) * invstd;
    float a = fmaxf(g * xh + be, 0.f);
    if (train) {
      Z[i] = xh;
      a = ctd_tr_philox_word(seed, step, layer, (uint32_t)i) < CTD_TR_KEEP_U32 ? a * 1.25f : 0.f;   // 1 / (1 - 0.2)
    }
    H[i] = a;
  }
}

// backward of the same: dH (gradient w.r.t. the dropout output) -> dZ (in place), dgamma, dbeta
//   h > 0 identifies both the kept mask and the open ReLU (a dropped or closed unit passes no gradient)
__global__ void __launch_bounds__(256) ctd_k_tr_bn_bwd(float* __restrict__ dH, const float* __restrict__ XH, const float* __restrict__ H, int N,
                                                       uint32_t B, uint32_t Bp, const float* __restrict__ gamma, const float* __restrict__ invstd,
                                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y, col = blockIdx.x * 32 + tx;
  float sg = 0.f, sb = 0.f;
  for (uint32_t r = ty; r < B; r += 8) {
    const size_t i = (size_t)r * N + col;
    const float da = H[i] > 0.f ? dH[i] * 1.25f : 0.f;
    dH[i] = da;
    sg += da * XH[i];
    sb += da;
  }
  red[ty][tx] = sg;
  __syncthreads();
  sg = 0.f;
  for (int i = 0; i < 8; ++i) sg += red[i][tx];
  __syncthreads();
  red[ty][tx] = sb;
  __syncthreads();
  sb = 0.f;
  for (int i = 0; i < 8; ++i) sb += red[i][tx];
  if (ty == 0) { dgamma[col] = sg; dbeta[col] = sb; }
  const float k = gamma[col] * invstd[col] / (float)B;
  for (uint32_t r = ty; r < Bp; r += 8) {
    const size_t i = (size_t)r * N + col;
    dH[i] = r < B ? k * ((float)B * dH[i] - sb - XH[i] * sg) : 0.f;
  }
}

// fc4 (128 -> 6), the loss and its gradient, one thread per row:
//   y = h3.W4^T + b4; p = y^2 / sum(y^2); loss_row = sum_j t_j (log t_j - log(p_j + 1e-10)), t = v^2 / sum(v^2)  (fp64, like the
//   reference: its labels are float64 and KLDivLoss promotes);  dL/dy_k = (2 y_k / S) (g_k - sum_j g_j p_j), g_j = -t_j / (B (p_j + 1e-10))
// loss_sum accumulates sum over rows (the caller divides by B: reduction='batchmean').
__global__ void __launch_bounds__(128) ctd_k_tr_head(const float* __restrict__ H3, const float* __restrict__ W4, const float* __restrict__ b4,
                                                     const double* __restrict__ T, uint32_t B, float* __restrict__ dY, double* __restrict__ loss_sum) {
  __shared__ float w[6 * CTD_TR_H3 + 6];
  __shared__ double part[128];
  for (int i = threadIdx.x; i < 6 * CTD_TR_H3; i += 128) w[i] = W4[i];
  if (threadIdx.x < 6) w[6 * CTD_TR_H3 + threadIdx.x] = b4[threadIdx.x];
  __syncthreads();
  const uint32_t r = blockIdx.x * 128 + threadIdx.x;
  double lrow = 0.0;
  if (r < B) {
    float y[6];
#pragma unroll
    for (int o = 0; o < 6; ++o) {
      float acc = 0.f;
      for (int k = 0; k < CTD_TR_H3; ++k) acc = fmaf(w[o * CTD_TR_H3 + k], H3[(size_t)r * CTD_TR_H3 + k], acc);
      y[o] = acc + w[6 * CTD_TR_H3 + o];
    }
    float S = 0.f, p[6];
#pragma unroll
    for (int o = 0; o < 6; ++o) { p[o] = y[o] * y[o]; S += p[o]; }
#pragma unroll
    for (int o = 0; o < 6; ++o) p[o] = p[o] / S;
    double ts = 0.0, t[6];
#pragma unroll
    for (int o = 0; o < 6; ++o) { t[o] = T[(size_t)r * 6 + o] * T[(size_t)r * 6 + o]; ts += t[o]; }
    double g[6], gp = 0.0;
#pragma unroll
    for (int o = 0; o < 6; ++o) {
      t[o] = t[o] / ts;
      const double lp = log((double)(p[o] + 1e-10f));
      if (t[o] > 0.0) lrow += t[o] * (log(t[o]) - lp);
      g[o] = -t[o] / ((double)B * (double)(p[o] + 1e-10f));
      gp += g[o] * (double)p[o];
    }
    if (dY != nullptr) {
#pragma unroll
      for (int o = 0; o < 6; ++o) dY[(size_t)r * 8 + o] = (float)((2.0 * (double)y[o] / (double)S) * (g[o] - gp));
    }
  }
  part[threadIdx.x] = lrow;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(loss_sum, part[0]);
}

// backward through fc4: dH3 [B][128] = dY.W4 masked by relu (h3 > 0); dW4 [6][128] = dY^T.H3; db4
__global__ void __launch_bounds__(128) ctd_k_tr_fc4_bwd(const float* __restrict__ dY, const float* __restrict__ W4, const float* __restrict__ H3,
                                                        uint32_t B, uint32_t Bp, float* __restrict__ dZ3) {
  const uint32_t r = blockIdx.x;
  const int k = threadIdx.x;
  float acc = 0.f;
  if (r < B) {
#pragma unroll
    for (int o = 0; o < 6; ++o) acc = fmaf(dY[(size_t)r * 8 + o], W4[o * CTD_TR_H3 + k], acc);
    if (!(H3[(size_t)r * CTD_TR_H3 + k] > 0.f)) acc = 0.f;
  }
  dZ3[(size_t)r * CTD_TR_H3 + k] = acc;   // rows B..Bp are zero
}
__global__ void __launch_bounds__(128) ctd_k_tr_fc4_wgrad(const float* __restrict__ dY, const float* __restrict__ H3, uint32_t B,
                                                          float* __restrict__ dW4, float* __restrict__ db4) {
  const int o = blockIdx.x, k = threadIdx.x;
  float acc = 0.f, bs = 0.f;
  for (uint32_t r = 0; r < B; ++r) {
    const float d = dY[(size_t)r * 8 + o];
    acc = fmaf(d, H3[(size_t)r * CTD_TR_H3 + k], acc);
    bs += d;
  }
  dW4[o * CTD_TR_H3 + k] = acc;
  if (k == 0) db4[o] = bs;
}
// column sums of a [B][N] gradient -> bias gradient
__global__ void __launch_bounds__(256) ctd_k_tr_colsum(const float* __restrict__ D, int N, uint32_t B, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y, col = blockIdx.x * 32 + tx;
  float s = 0.f;
  for (uint32_t r = ty; r < B; r += 8) s += D[(size_t)r * N + col];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    out[col] = s;
  }
}
// relu mask for fc3's output gradient is folded into ctd_k_tr_fc4_bwd; relu on fc3's forward is done by the GEMM epilogue

// torch.optim.Adam (betas 0.9 / 0.999, eps 1e-8, no weight decay, no amsgrad), bias-corrected
__global__ void ctd_k_tr_adam(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ M, float* __restrict__ V, size_t n,
                              float lr, float bc1, float bc2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = G[i];
  const float m = 0.9f * M[i] + 0.1f * g;
  const float v = 0.999f * V[i] + 0.001f * g * g;
  M[i] = m; V[i] = v;
  const float denom = sqrtf(v) / sqrtf(bc2) + 1e-8f;
  P[i] -= (lr / bc1) * (m / denom);
}
