// Weight gradient of 3x3 / pad 1 convolutions with 64 input channels (the widest-resolution layers of BigGAN-Deep: the
// generic wgrad kernel reloads dY and a shifted X tile for every one of the 9 taps and wastes half of each M = 128 MMA
// when Cout = 64; it runs at 235 TFLOP/s, bound by L2 -> smem traffic).
//
//   dW[co][tap][ci] = sum_{b,h,w} dY[b,h,w,co] * X[b,h+dh,w+dw,ci]
//
// Formulation (pixels are the contraction dimension, all operands MN-major / SWIZZLE_128B):
//   A (M side) = X halo: one tile stages 4 halo rows x 130 pixels x 64 channels ONCE; tap (dh,dw) is the same buffer read
//                from pixel offset (dh+1)*130 + (dw+1).  An M = 128 operand consists of two 64-channel atoms "LBO" bytes
//                apart -- choosing LBO = the pixel distance between two taps makes ONE MMA produce TWO taps:
//                D[(tap_pair, ci)][co].  9 taps = 4 pairs + 1 single -> 5 accumulator blocks of 64 columns (320 of 512).
//   B (N side) = dY tile, 2 rows x 128 pixels x Cout (<= 64) channels.
// Per tile: 2 rows x 5 blocks x 8 MMAs (K = 16 pixels) of 128x64x16 = 2560 tensor cycles for 98.5 KB of loads
// (38 B/cycle/SM), 90 % of the MMAs' rows useful.  A CTA accumulates its whole pixel range in TMEM and drains once with
// coalesced fp32 red.adds (lanes = consecutive ci).
#include "common.cuh"
#include "ptx.cuh"

namespace sgb {

static constexpr int kW3Threads = 192;              // warp 0 producer, warp 1 MMA + TMEM owner, warps 2-5 drain
static constexpr int kW3Px = 128;                   // pixels per tile row
static constexpr int kW3Halo = kW3Px + 2;           // halo row pitch in pixels
static constexpr uint32_t kW3DyBytes = 2 * kW3Px * 128;                       // 32 KB
static constexpr uint32_t kW3XBytes = 4 * kW3Halo * 128;                      // 66 560 B
static constexpr uint32_t kW3XAlloc = (kW3XBytes + 2048 + 1023) / 1024 * 1024;  // + slack: the unpaired tap's idle atom reads past
static constexpr uint32_t kW3Stage = kW3DyBytes + kW3XAlloc;
static constexpr int kW3Stages = 2;

struct W3Args {
  int H, W, Cin, Cout;           // Cin: channels of THIS launch's 64-wide input-channel block
  int dw_cin;                    // row pitch of dw (= the layer's full Cin)
  int tiles_w, tiles_h;          // tiles per row, row pairs per image
  long long tiles;               // B * tiles_h * tiles_w
  float* dw;
  float* dbias;                  // optional: sum of dY over all pixels (idle atom of the unpaired tap reads constant ones)
};

__global__ void __launch_bounds__(kW3Threads, 1)
wgrad3x3_c64_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const W3Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kW3Stages * kW3Stage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kW3Stages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kW3Stages);
  const uint32_t holder = bar_base + 8u * (2 * kW3Stages + 1);
  const uint32_t ones_base = (bar_base + 8u * (2 * kW3Stages + 2) + 1023u) & ~1023u;   // 2 KB of bf16 1.0 (16 K rows x 128 B)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kW3Stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(holder, 512);
    tmem_relinquish();
  }
  if (p.dbias) {
    for (int i = threadIdx.x; i < 2048 / 4; i += kW3Threads)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(ones_base + 4u * i), "r"(0x3F803F80u) : "memory");
    fence_proxy_async_smem();                     // the MMA reads this region through the async proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(holder));

  // contiguous tile range of this CTA
  const long long t_begin = p.tiles * blockIdx.x / gridDim.x;
  const long long t_end = p.tiles * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const int wt = (int)(t % p.tiles_w);
        const int ht = (int)((t / p.tiles_w) % p.tiles_h);
        const int b = (int)(t / ((long long)p.tiles_w * p.tiles_h));
        const int w0 = wt * kW3Px, h0 = ht * 2;
        const int s = it % kW3Stages;
        const uint32_t ph = (it / kW3Stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), kW3DyBytes + kW3XBytes);
        const uint32_t sa = smem_base + s * kW3Stage;
        tma_load_4d(sa, &tmDY, full_bar(s), 0, w0, h0, b);
        tma_load_4d(sa + kW3DyBytes, &tmX, full_bar(s), 0, w0 - 1, h0 - 1, b);   // out-of-range halo = zero padding
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && t_end > t_begin) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);   // both operands MN-major
      uint32_t it = 0;
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const int s = it % kW3Stages;
        const uint32_t ph = (it / kW3Stages) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t ys = smem_base + s * kW3Stage;
        const uint32_t xs = ys + kW3DyBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int g = 0; g < 5; ++g) {
            const int tap = 2 * g;
            const int off = (tap / 3 + j) * kW3Halo + tap % 3;                 // pixel offset of the pair's first tap
            // second atom = next tap: +1 pixel, or (row + 1, dw = -1) for the pair (tap 2, tap 3): 130 - 2 pixels;
            // the unpaired tap 8 drags an idle atom one pixel further (rows ignored by the drain, reads stay in the slack)
            const uint32_t lbo = (tap == 2 ? (uint32_t)(kW3Halo - 2) : 1u) * 128u;
            const uint32_t a0 = xs + (uint32_t)off * 128u;
            const uint64_t adesc = make_sdesc_sw128(a0, lbo, 1024);
            const uint64_t bdesc = make_sdesc_sw128(ys + (uint32_t)(j * kW3Px) * 128u, 128 * kW3Px, 1024);
            const uint32_t d_tmem = tmem_base + (uint32_t)g * 64u;
#pragma unroll
            for (int kk = 0; kk < kW3Px / 16; ++kk) {
              // 16 pixels (K) = two 8-row groups = 2048 bytes -> +128 in the (addr >> 4) field
              uint64_t ad = adesc + 128 * kk;
              // bias gradient for free: the idle second atom of the unpaired tap 8 is pointed (through the LBO) at a
              // constant-one K slab, so rows 64..127 of its accumulator block receive sum_k dY[k][co]
              if (g == 4 && p.dbias) ad = make_sdesc_sw128(a0 + 2048u * kk, ones_base - (a0 + 2048u * kk), 1024);
              umma_f16_ss(d_tmem, ad, bdesc + 128 * kk, idesc, (it > 0 || j > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
  } else if (t_end > t_begin) {
    const int q = warp & 3;
    const int m = q * 32 + lane;                 // accumulator row = (tap parity, ci)
    const int ci = m & 63;
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int g = 0; g < 5; ++g) {
      const int tap = 2 * g + (m >> 6);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 64u;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(t_row + c0, v);
        tmem_ld_wait();
        if (tap == 9 && m == 64 && p.dbias) {          // row (tap 9, ci 0): the bias gradient
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < p.Cout) atomicAdd(p.dbias + c0 + i, __uint_as_float(v[i]));
        }
        if (tap >= 9 || ci >= p.Cin) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int co = c0 + i;
          if (co < p.Cout) atomicAdd(p.dw + ((long long)co * 9 + tap) * p.dw_cin + ci, __uint_as_float(v[i]));
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool wgrad3x3_c64_eligible(const sgb_wgrad_desc* d) {
  // Cin > 64 (the generator's 128 -> 3 output convolution) runs as one launch per 64-channel block of the input
  return d->KH == 3 && d->KW == 3 && d->pad_h == 1 && d->pad_w == 1 && !d->per_image &&
         ((d->Cin <= 64 && d->Cin % 8 == 0) || (d->Cin % 64 == 0 && d->Cin <= 256 && d->Cout <= 16)) &&
         d->Cout <= 64 && d->Cout % 8 == 0 && d->W % kW3Px == 0 && d->H % 2 == 0 &&
         (long long)d->B * (d->H / 2) * (d->W / kW3Px) >= 64;
}

int launch_wgrad3x3_c64(const sgb_wgrad_desc* d, cudaStream_t stream) {
  if (!d->accumulate) {
    SGB_CUDA(cudaMemsetAsync(d->dw, 0, sizeof(float) * (size_t)d->Cout * 9 * d->Cin, stream));
    if (d->dbias) SGB_CUDA(cudaMemsetAsync(d->dbias, 0, sizeof(float) * (size_t)d->Cout, stream));
  }
  CUtensorMap tmDY;
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->dy_cstride * 2, (uint64_t)d->dy_cstride * 2 * d->W, (uint64_t)d->dy_cstride * 2 * d->W * d->H};
    uint32_t box[4] = {64, (uint32_t)kW3Px, 2, 1};
    int rc = make_tmap_bf16(&tmDY, d->dy, 4, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = (size_t)kW3Stages * kW3Stage + 1024 + 8 * (2 * kW3Stages + 2) + 1024 + 2048 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    SGB_CUDA(cudaFuncSetAttribute(wgrad3x3_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  for (int ci0 = 0; ci0 < d->Cin; ci0 += 64) {
    W3Args p;
    p.H = d->H; p.W = d->W; p.Cout = d->Cout;
    p.Cin = d->Cin - ci0 < 64 ? d->Cin - ci0 : 64;
    p.dw_cin = d->Cin;
    p.tiles_w = d->W / kW3Px;
    p.tiles_h = d->H / 2;
    p.tiles = (long long)d->B * p.tiles_h * p.tiles_w;
    p.dw = d->dw + ci0;
    p.dbias = ci0 == 0 ? d->dbias : nullptr;
    CUtensorMap tmX;
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t strides[3] = {(uint64_t)d->x_cstride * 2, (uint64_t)d->x_cstride * 2 * d->W, (uint64_t)d->x_cstride * 2 * d->W * d->H};
    uint32_t box[4] = {64, (uint32_t)kW3Halo, 4, 1};
    int rc = make_tmap_bf16(&tmX, (const bf16*)d->x + ci0, 4, dims, strides, box);
    if (rc) return rc;
    const int grid = p.tiles < sm_count() ? (int)p.tiles : sm_count();
    wgrad3x3_c64_kernel<<<grid, kW3Threads, smem, stream>>>(tmDY, tmX, p);
    SGB_LAUNCH_CHECK();
  }
  return SGB_OK;
}

}  // namespace sgb
