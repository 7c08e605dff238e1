// Spectral normalisation: one power iteration, sigma, and emission of the tensor-core weight packs.
//
// Reference arithmetic (torch/nn/utils/spectral_norm.py:62-114, used by src/utils/ops.py:195-224 with eps=1e-6,
// n_power_iterations=1, dim=0):   v <- normalize(W^T u);  u <- normalize(W v);  sigma = u . (W v);  W_sn = W / sigma
// where W is the weight viewed as [R = out_channels, K = in_channels*kh*kw] and normalize(x) = x / max(||x||_2, eps).
// HBM-bound: two streaming passes over the fp32 weight for the iteration and one for the packs.
#include "common.cuh"

namespace sgb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  r = sh[0];
  return r;
}

// ws layout (floats): [0..K) t = W^T u accumulator, [K..K+R) s = W v, [K+R] ticket A, [K+R+1] ticket B
// Pass 1: t += W^T u over a (rows_per_block x 256-column) tile; the last block normalises t into v and clears t.
__global__ void __launch_bounds__(256) sn_wtu_kernel(const float* __restrict__ W, const float* __restrict__ u,
                                                      float* __restrict__ v, float* __restrict__ ws, int R, int K,
                                                      int rows_per_block, float eps) {
  __shared__ float sh[32];
  __shared__ bool last;
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, R);
  if (c < K) {
    float acc = 0.f;
    const float* wp = W + (size_t)r0 * K + c;
    for (int r = r0; r < r1; ++r, wp += K) acc = fmaf(__ldg(wp), __ldg(u + r), acc);
    atomicAdd(ws + c, acc);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = reinterpret_cast<unsigned*>(ws + K + R);
    const unsigned total = gridDim.x * gridDim.y;
    last = (atomicAdd(ticket, 1u) == total - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float ss = 0.f;
  for (int i = threadIdx.x; i < K; i += 256) { const float t = __ldcg(ws + i); ss = fmaf(t, t, ss); }
  ss = block_sum(ss, sh);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  for (int i = threadIdx.x; i < K; i += 256) {
    v[i] = __ldcg(ws + i) * inv;
    ws[i] = 0.f;
  }
}

// Pass 2: s = W v (one warp per row); the last block forms u = normalize(s), sigma = u . s.
__global__ void __launch_bounds__(256) sn_wv_kernel(const float* __restrict__ W, float* __restrict__ u,
                                                     const float* __restrict__ v, float* __restrict__ ws,
                                                     float* __restrict__ sigma, int R, int K, float eps, int update_u) {
  __shared__ float sh[32];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s = ws + K;
  for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
    const float* wp = W + (size_t)r * K;
    float acc = 0.f;
    for (int i = lane; i < K; i += 32) acc = fmaf(__ldg(wp + i), __ldg(v + i), acc);
    acc = warp_sum(acc);
    if (lane == 0) s[r] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = reinterpret_cast<unsigned*>(ws + K + R + 1);
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (update_u) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < R; i += 256) { const float t = __ldcg(s + i); ss = fmaf(t, t, ss); }
    ss = block_sum(ss, sh);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    for (int i = threadIdx.x; i < R; i += 256) u[i] = __ldcg(s + i) * inv;
    if (threadIdx.x == 0) *sigma = ss * inv;  // u . s = ||s||^2 / max(||s||, eps)
  } else {
    float d = 0.f;
    for (int i = threadIdx.x; i < R; i += 256) d = fmaf(__ldcg(s + i), u[i], d);
    d = block_sum(d, sh);
    if (threadIdx.x == 0) *sigma = d;
  }
}

// Packs: W is the module's weight [R = Cout][Cin][taps] (taps = kh*kw, row-major OIHW).
//   fprop pack  Wf[co][tap][ci]            = W[co][ci][tap] / sigma
//   dgrad pack  Wd[ci][taps-1-tap][co]     = W[co][ci][tap] / sigma     (180-degree rotated, in/out swapped)
// rows_perm > 1 re-orders output rows for linear0, whose output is viewed as [C, S] by the reference
// (src/models/big_resnet_deep_legacy.py:167-168) but consumed here as NHWC [S, C]: packed row s*C + c <- row c*S + s.
__global__ void __launch_bounds__(256) sn_pack_kernel(const float* __restrict__ W, const float* __restrict__ sigma,
                                                       bf16* __restrict__ wf, bf16* __restrict__ wd, int Cout, int Cin,
                                                       int taps, int perm_S, int Cout_p, int Cin_p) {
  const size_t total = (size_t)Cout * Cin * taps;
  const float inv = sigma ? 1.f / __ldg(sigma) : 1.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    int co = (int)(i / ((size_t)taps * Cin));
    const float val = __ldg(W + i) * inv;
    if (perm_S > 1) {
      const int C = Cout / perm_S;
      co = (co % perm_S) * C + co / perm_S;
    }
    const bf16 b = __float2bfloat16_rn(val);
    if (wf) wf[((size_t)co * taps + tap) * Cin_p + ci] = b;
    if (wd) wd[((size_t)ci * taps + (taps - 1 - tap)) * Cout_p + co] = b;
  }
}

// Backward of W_sn = W / sigma with u, v held constant (they are detached buffers in the reference):
//   dL/dW = (G - <G, W_sn> u v^T) / sigma,  G given in the fprop pack layout [co][tap][ci] (fp32, from wgrad).
__global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const float* __restrict__ G, const float* __restrict__ W,
                                                          float* __restrict__ dot, int Cout, int Cin, int taps, int perm_S,
                                                          int Cin_p) {
  __shared__ float sh[32];
  const size_t total = (size_t)Cout * Cin * taps;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    int co = (int)(i / ((size_t)taps * Cin));
    if (perm_S > 1) { const int C = Cout / perm_S; co = (co % perm_S) * C + co / perm_S; }
    acc = fmaf(__ldg(G + ((size_t)co * taps + tap) * Cin_p + ci), __ldg(W + i), acc);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(dot, acc);
}

__global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const float* __restrict__ G, const float* __restrict__ u,
                                                            const float* __restrict__ v, const float* __restrict__ sigma,
                                                            const float* __restrict__ dot, float* __restrict__ dW, int Cout,
                                                            int Cin, int taps, int perm_S, int accumulate, int Cin_p) {
  const size_t total = (size_t)Cout * Cin * taps;
  float inv = 1.f, coef = 0.f;
  if (sigma) {
    inv = 1.f / __ldg(sigma);
    coef = __ldg(dot) * inv;  // <G, W> / sigma = <G, W_sn>
  }
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co_orig = (int)(i / ((size_t)taps * Cin));
    int co = co_orig;
    if (perm_S > 1) { const int C = Cout / perm_S; co = (co % perm_S) * C + co / perm_S; }
    float g = __ldg(G + ((size_t)co * taps + tap) * Cin_p + ci);
    if (sigma) g = (g - coef * __ldg(u + co_orig) * __ldg(v + (size_t)ci * taps + tap)) * inv;
    dW[i] = accumulate ? dW[i] + g : g;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Batched variant: every spectrally-normalised layer of a network in three launches (blockIdx.y = layer).  The per-layer
// launches above cost ~30 us each in dependent tiny kernels; a BigGAN-Deep generator has 150+ such layers per forward.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sn_wtu_batch_kernel(const sgb_sn_layer* __restrict__ table, float eps) {
  __shared__ float sh[32];
  __shared__ bool last;
  const sgb_sn_layer L = table[blockIdx.y];
  if (!L.has_sn) return;
  const int R = L.R, K = L.K;
  const int col_blocks = (K + 255) / 256, rows_per_block = 64;
  const int row_blocks = (R + rows_per_block - 1) / rows_per_block;
  const unsigned total = (unsigned)col_blocks * row_blocks;
  if (blockIdx.x >= total) return;
  const int bx = blockIdx.x % col_blocks, by = blockIdx.x / col_blocks;
  const float* W = L.W;
  float* ws = L.ws;
  const int c = bx * 256 + threadIdx.x;
  const int r0 = by * rows_per_block, r1 = min(r0 + rows_per_block, R);
  if (c < K) {
    float acc = 0.f;
    const float* wp = W + (size_t)r0 * K + c;
    for (int r = r0; r < r1; ++r, wp += K) acc = fmaf(__ldg(wp), __ldg(L.u + r), acc);
    atomicAdd(ws + c, acc);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = reinterpret_cast<unsigned*>(ws + K + R);
    last = (atomicAdd(ticket, 1u) == total - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float ss = 0.f;
  for (int i = threadIdx.x; i < K; i += 256) { const float t = __ldcg(ws + i); ss = fmaf(t, t, ss); }
  ss = block_sum(ss, sh);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  for (int i = threadIdx.x; i < K; i += 256) {
    L.v[i] = __ldcg(ws + i) * inv;
    ws[i] = 0.f;
  }
}

__global__ void __launch_bounds__(256) sn_wv_batch_kernel(const sgb_sn_layer* __restrict__ table, float* __restrict__ sigma_all,
                                                           float eps, int update_u) {
  __shared__ float sh[32];
  __shared__ bool last;
  const sgb_sn_layer L = table[blockIdx.y];
  if (!L.has_sn) {
    if (blockIdx.x == 0 && threadIdx.x == 0) sigma_all[blockIdx.y] = 1.f;
    return;
  }
  const int R = L.R, K = L.K;
  const unsigned nblk = min((unsigned)gridDim.x, (unsigned)((R + 7) / 8));
  if (blockIdx.x >= nblk) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s = L.ws + K;
  for (int r = blockIdx.x * 8 + warp; r < R; r += nblk * 8) {
    const float* wp = L.W + (size_t)r * K;
    float acc = 0.f;
    for (int i = lane; i < K; i += 32) acc = fmaf(__ldg(wp + i), __ldcg(L.v + i), acc);
    acc = warp_sum(acc);
    if (lane == 0) s[r] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = reinterpret_cast<unsigned*>(L.ws + K + R + 1);
    last = (atomicAdd(ticket, 1u) == nblk - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float* sigma = sigma_all + blockIdx.y;
  if (update_u) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < R; i += 256) { const float t = __ldcg(s + i); ss = fmaf(t, t, ss); }
    ss = block_sum(ss, sh);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    for (int i = threadIdx.x; i < R; i += 256) L.u[i] = __ldcg(s + i) * inv;
    if (threadIdx.x == 0) *sigma = ss * inv;
  } else {
    float d = 0.f;
    for (int i = threadIdx.x; i < R; i += 256) d = fmaf(__ldcg(s + i), L.u[i], d);
    d = block_sum(d, sh);
    if (threadIdx.x == 0) *sigma = d;
  }
}

__global__ void __launch_bounds__(256) sn_pack_batch_kernel(const sgb_sn_layer* __restrict__ table, int n_layers, int total_tiles,
                                                             const float* __restrict__ sigma_all, bf16* __restrict__ pack_f,
                                                             bf16* __restrict__ pack_d) {
  // Tile = 32 output channels x 32 input channels x all taps, staged in shared memory so that the fp32 reads (one run of
  // 32*taps floats per output channel) AND both bf16 writes are contiguous: the fprop pack [co][tap][ci] in runs of 32 ci,
  // the dgrad pack [ci][taps-1-tap][co] in runs of 32 co.  (The element-wise version scattered 2-byte stores: 0.5 ms/call.)
  // Work decomposition: ONE global tile index over all layers (table[l].tile_start = tiles of the layers before l; layers
  // with more than 9 taps count 4096-element pseudo tiles); a block owns a contiguous range of it.  The former
  // (max tiles) x (layers) grid launched ~300 k blocks of which most returned at once: 0.49 ms per call (r02 launch list).
  constexpr int kT = 32, kMaxTaps = 9;
  __shared__ bf16 tile[kT][kT * kMaxTaps + 2];
  // (total_tiles / tile_start count WORK UNITS of ~1024 weights: a tile of a k-tap layer is k units, a pseudo tile 4, so
  //  that contiguous unit ranges are balanced; a tile belongs to the block whose range holds its first unit)
  const int u0 = (int)((long long)total_tiles * blockIdx.x / gridDim.x);
  const int u1 = (int)((long long)total_tiles * (blockIdx.x + 1) / gridDim.x);
  if (u0 >= u1) return;
  int lo = 0, hi = n_layers - 1;
  while (lo < hi) {                                // last layer whose first unit is <= u0
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].tile_start <= u0) lo = mid; else hi = mid - 1;
  }
  int l = lo;
  sgb_sn_layer L = table[l];
  int next_start = (l + 1 < n_layers) ? table[l + 1].tile_start : total_tiles;
  float inv = 1.f / __ldcg(sigma_all + l);
  int upt = L.taps > kMaxTaps ? 4 : L.taps;        // units per tile of the current layer
  int t = (u0 - L.tile_start + upt - 1) / upt;     // first tile starting at or after u0
  for (;;) {
    int ustart = L.tile_start + t * upt;
    while (ustart >= next_start && l + 1 < n_layers) {   // next layer (layers without tiles are skipped here)
      ++l;
      L = table[l];
      next_start = (l + 1 < n_layers) ? table[l + 1].tile_start : total_tiles;
      inv = 1.f / __ldcg(sigma_all + l);
      upt = L.taps > kMaxTaps ? 4 : L.taps;
      t = 0;
      ustart = L.tile_start;
    }
    if (ustart >= u1 || ustart >= next_start) break;
    const int Cout = L.Cout, Cin = L.Cin, taps = L.taps, perm_S = L.perm_S;
    bf16* wf = pack_f ? pack_f + L.off_f : nullptr;
    bf16* wd = pack_d ? pack_d + L.off_d : nullptr;
    if (taps > kMaxTaps) {                          // element-wise pseudo tile (4x4 filters of the DCGAN generator)
      const unsigned total = (unsigned)Cout * Cin * taps;
      const unsigned e0 = (unsigned)t * 4096u, e1 = min(e0 + 4096u, total);
      for (unsigned i = e0 + threadIdx.x; i < e1; i += 256) {
        const int tap = (int)(i % taps);
        const int ci = (int)((i / taps) % Cin);
        int co = (int)(i / ((unsigned)taps * Cin));
        const bf16 v = __float2bfloat16_rn(__ldg(L.W + i) * inv);
        if (perm_S > 1) { const int C = Cout / perm_S; co = (co % perm_S) * C + co / perm_S; }
        if (wf) wf[((size_t)co * taps + tap) * L.Cin_p + ci] = v;
        if (wd) wd[((size_t)ci * taps + (taps - 1 - tap)) * L.Cout_p + co] = v;
      }
      ++t;
      continue;
    }
    const int tiles_ci = (Cin + kT - 1) / kT;
    const int run = kT * taps;                      // source elements per output channel inside one tile
    const int co0 = (t / tiles_ci) * kT, ci0 = (t % tiles_ci) * kT;
    const int nci = min(kT, Cin - ci0), nco = min(kT, Cout - co0);
    __syncthreads();
    // read: row r of the tile = source channel co0 + r, elements (ci0*taps .. (ci0+nci)*taps) contiguous in memory
    for (int e = threadIdx.x; e < kT * run; e += 256) {
      const int r = e / run, k = e % run;
      if (r < nco && k < nci * taps)
        tile[r][k] = __float2bfloat16_rn(__ldg(L.W + ((size_t)(co0 + r) * Cin + ci0) * taps + k) * inv);
    }
    __syncthreads();
    if (wf) {   // [co][tap][ci]: runs of nci consecutive ci
      for (int e = threadIdx.x; e < kT * run; e += 256) {
        const int r = e / run, rem = e % run;
        const int tap = rem / kT, c = rem % kT;
        if (r < nco && c < nci) {
          int co = co0 + r;
          if (perm_S > 1) { const int C = Cout / perm_S; co = (co % perm_S) * C + co / perm_S; }
          wf[((size_t)co * taps + tap) * L.Cin_p + ci0 + c] = tile[r][c * taps + tap];
        }
      }
    }
    if (wd) {   // [ci][taps-1-tap][co]: runs of nco consecutive co (perm_S > 1 only permutes the rows of linear0: scattered there)
      for (int e = threadIdx.x; e < kT * run; e += 256) {
        const int k = e / kT, r = e % kT;          // k = c * taps + tap
        const int c = k / taps, tap = k % taps;
        if (r < nco && c < nci) {
          int co = co0 + r;
          if (perm_S > 1) { const int C = Cout / perm_S; co = (co % perm_S) * C + co / perm_S; }
          wd[((size_t)(ci0 + c) * taps + (taps - 1 - tap)) * L.Cout_p + co] = tile[r][k];
        }
      }
    }
    ++t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Batched spectral-norm backward: dW += (G - <G, W/sigma> u v^T) / sigma for EVERY layer of a network pass in two launches
// (blockIdx.y = layer).  G lives in one flat fp32 buffer laid out like the fprop packs (layer l at g_flat + off_f[l]); u / v /
// sigma are the copies taken at that forward pass (flat arenas, per-layer offsets in the table).  Layers that produced no
// weight gradient in this pass hold zeros in g_flat and add zero.
// ---------------------------------------------------------------------------------------------------------------------
// Both kernels walk a layer row by row (row = one output channel: Cin * taps contiguous floats of W / dW, and the
// contiguous [taps][Cin_p] block of G).  For taps > 1 the G row is staged in shared memory so that the global reads of G, W and
// the read-modify-write of dW are all coalesced (the element-wise version read G with a stride of Cin_p floats and did three
// 64-bit divisions per element: 0.25 ms per kernel and call).
static constexpr int kSnRowFloats = 8192;          // 32 KiB: 512 channels x 16 taps

__device__ __forceinline__ int sn_g_row(const sgb_snbwd_layer& L, int co_orig) {
  if (L.perm_S > 1) { const int C = L.Cout / L.perm_S; return (co_orig % L.perm_S) * C + co_orig / L.perm_S; }
  return co_orig;
}

__global__ void __launch_bounds__(256) sn_bwd_dot_batch_kernel(const sgb_snbwd_layer* __restrict__ table,
                                                                const float* __restrict__ g_flat, float* __restrict__ dots) {
  __shared__ float sh[32];
  __shared__ float grow[kSnRowFloats];
  const sgb_snbwd_layer L = table[blockIdx.y];
  if (!L.has_sn || !L.dW) return;
  const int taps = L.taps, Cin = L.Cin, K = Cin * taps, GK = taps * L.Cin_p;
  const bool staged = taps > 1 && GK <= kSnRowFloats;
  float acc = 0.f;
  for (int r = blockIdx.x; r < L.Cout; r += gridDim.x) {
    const float* G = g_flat + L.off_g + (size_t)sn_g_row(L, r) * GK;
    const float* W = L.W + (size_t)r * K;
    if (taps == 1) {
      for (int k = threadIdx.x; k < K; k += 256) acc = fmaf(__ldg(G + k), __ldg(W + k), acc);
    } else if (staged) {
      __syncthreads();
      for (int k = threadIdx.x; k < GK; k += 256) grow[k] = __ldg(G + k);
      __syncthreads();
      for (int k = threadIdx.x; k < K; k += 256) {
        const int ci = k / taps, tap = k - ci * taps;
        acc = fmaf(grow[tap * L.Cin_p + ci], __ldg(W + k), acc);
      }
    } else {
      for (int k = threadIdx.x; k < K; k += 256) {
        const int ci = k / taps, tap = k - ci * taps;
        acc = fmaf(__ldg(G + (size_t)tap * L.Cin_p + ci), __ldg(W + k), acc);
      }
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0 && acc != 0.f) atomicAdd(dots + blockIdx.y, acc);
}

__global__ void __launch_bounds__(256) sn_bwd_apply_batch_kernel(const sgb_snbwd_layer* __restrict__ table,
                                                                  const float* __restrict__ g_flat, const float* __restrict__ dots,
                                                                  const float* __restrict__ sigma_all, const float* __restrict__ u_flat,
                                                                  const float* __restrict__ v_flat) {
  __shared__ float grow[kSnRowFloats];
  const sgb_snbwd_layer L = table[blockIdx.y];
  if (!L.dW) return;
  const int taps = L.taps, Cin = L.Cin, K = Cin * taps, GK = taps * L.Cin_p;
  const bool staged = taps > 1 && GK <= kSnRowFloats;
  float inv = 1.f, coef = 0.f;
  const float* u = nullptr;
  const float* v = nullptr;
  if (L.has_sn) {
    inv = 1.f / __ldg(sigma_all + blockIdx.y);
    coef = __ldg(dots + blockIdx.y) * inv;
    u = u_flat + L.off_u;
    v = v_flat + L.off_v;
  }
  for (int r = blockIdx.x; r < L.Cout; r += gridDim.x) {
    const float* G = g_flat + L.off_g + (size_t)sn_g_row(L, r) * GK;
    float* dW = L.dW + (size_t)r * K;
    const float cu = L.has_sn ? coef * __ldg(u + r) : 0.f;
    if (staged) {
      __syncthreads();
      for (int k = threadIdx.x; k < GK; k += 256) grow[k] = __ldg(G + k);
      __syncthreads();
    }
    for (int k = threadIdx.x; k < K; k += 256) {
      float g;
      if (taps == 1) {
        g = __ldg(G + k);
      } else {
        const int ci = k / taps, tap = k - ci * taps;
        g = staged ? grow[tap * L.Cin_p + ci] : __ldg(G + (size_t)tap * L.Cin_p + ci);
      }
      if (L.has_sn) g = (g - cu * __ldg(v + k)) * inv;
      dW[k] += g;
    }
  }
}

}  // namespace sgb

using namespace sgb;

extern "C" int sgb_sn_backward_batch(const sgb_snbwd_layer* table, int32_t n_layers, const float* g_flat, const float* sigma_all,
                                     const float* u_flat, const float* v_flat, float* dots, int32_t max_blocks,
                                     sgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SGB_REQUIRE(table && n_layers > 0 && g_flat && dots && max_blocks > 0);
  SGB_CUDA(cudaMemsetAsync(dots, 0, sizeof(float) * (size_t)n_layers, stream));
  sn_bwd_dot_batch_kernel<<<dim3(max_blocks, n_layers), 256, 0, stream>>>(table, g_flat, dots);
  SGB_LAUNCH_CHECK();
  sn_bwd_apply_batch_kernel<<<dim3(max_blocks, n_layers), 256, 0, stream>>>(table, g_flat, dots, sigma_all, u_flat, v_flat);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

extern "C" int sgb_sn_batch(const sgb_sn_layer* table, int32_t n_layers, float* sigma_all, void* pack_f, void* pack_d, float eps,
                            int32_t do_power_iteration, int32_t max_blocks_wtu, int32_t max_blocks_wv, int32_t max_blocks_pack,
                            sgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SGB_REQUIRE(table && n_layers > 0 && sigma_all && max_blocks_wtu > 0 && max_blocks_wv > 0 && max_blocks_pack > 0);
  if (do_power_iteration) {
    sn_wtu_batch_kernel<<<dim3(max_blocks_wtu, n_layers), 256, 0, stream>>>(table, eps);
    SGB_LAUNCH_CHECK();
  }
  sn_wv_batch_kernel<<<dim3(max_blocks_wv, n_layers), 256, 0, stream>>>(table, sigma_all, eps, do_power_iteration);
  SGB_LAUNCH_CHECK();
  if (pack_f || pack_d) {
    // max_blocks_pack carries the total tile count of the table (sum over layers, see sgb_sn_layer.tile_start)
    const int total_tiles = max_blocks_pack;
    const int blocks = total_tiles < 32 * sm_count() ? total_tiles : 32 * sm_count();   // 4 waves of resident blocks: dynamic balance
    sn_pack_batch_kernel<<<blocks, 256, 0, stream>>>(table, n_layers, total_tiles, sigma_all, (bf16*)pack_f, (bf16*)pack_d);
    SGB_LAUNCH_CHECK();
  }
  return SGB_OK;
}

extern "C" int64_t sgb_sn_workspace_floats(int32_t R, int32_t K) { return (int64_t)R + K + 4; }

extern "C" int sgb_sn_power_iter(const float* W, float* u, float* v, float* sigma, float* ws, int32_t R, int32_t K,
                                 float eps, int32_t do_power_iteration, sgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SGB_REQUIRE(W && u && v && sigma && ws && R > 0 && K > 0);
  if (do_power_iteration) {
    int rows_per_block = 64;
    dim3 grid((K + 255) / 256, (R + rows_per_block - 1) / rows_per_block);
    sn_wtu_kernel<<<grid, 256, 0, stream>>>(W, u, v, ws, R, K, rows_per_block, eps);
    SGB_LAUNCH_CHECK();
  }
  int blocks = (R + 7) / 8;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  sn_wv_kernel<<<blocks, 256, 0, stream>>>(W, u, v, ws, sigma, R, K, eps, do_power_iteration);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

extern "C" int sgb_weight_pack(const float* W, const float* sigma, void* w_fprop, void* w_dgrad, int32_t Cout, int32_t Cin,
                               int32_t taps, int32_t perm_S, int32_t Cout_p, int32_t Cin_p, sgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SGB_REQUIRE(W && (w_fprop || w_dgrad) && Cout > 0 && Cin > 0 && taps > 0);
  SGB_REQUIRE(perm_S <= 1 || Cout % perm_S == 0);
  SGB_REQUIRE(Cout_p >= Cout && Cin_p >= Cin);
  const size_t total = (size_t)Cout * Cin * taps;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  sn_pack_kernel<<<blocks, 256, 0, stream>>>(W, sigma, (bf16*)w_fprop, (bf16*)w_dgrad, Cout, Cin, taps, perm_S, Cout_p, Cin_p);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

extern "C" int sgb_sn_backward(const float* G, const float* W, const float* u, const float* v, const float* sigma,
                               float* scratch_dot, float* dW, int32_t Cout, int32_t Cin, int32_t taps, int32_t perm_S,
                               int32_t Cin_p, int32_t accumulate, sgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SGB_REQUIRE(G && dW && Cout > 0 && Cin > 0 && taps > 0);
  SGB_REQUIRE(!sigma || (W && u && v && scratch_dot));
  SGB_REQUIRE(Cin_p >= Cin);
  const size_t total = (size_t)Cout * Cin * taps;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  if (sigma) {
    SGB_CUDA(cudaMemsetAsync(scratch_dot, 0, sizeof(float), stream));
    sn_bwd_dot_kernel<<<blocks, 256, 0, stream>>>(G, W, scratch_dot, Cout, Cin, taps, perm_S, Cin_p);
    SGB_LAUNCH_CHECK();
  }
  sn_bwd_apply_kernel<<<blocks, 256, 0, stream>>>(G, u, v, sigma, scratch_dot, dW, Cout, Cin, taps, perm_S, accumulate, Cin_p);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
