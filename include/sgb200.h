/* libsgb200 — C ABI of the B200-native StudioGAN hot path.
 *
 * Everything here is plain C: device pointers, sizes, a cudaStream_t passed as void*.  No torch
 * types cross this boundary.  Every entry point returns 0 on success or an SGB_ERR_* code; nothing
 * throws, nothing allocates behind the caller's back unless the name says "workspace".
 *
 * The reference (POSTECH-CVLab/PyTorch-StudioGAN) has no FFI: its hot path is Python calling
 * torch/cuDNN/cuBLAS ops.  Each entry point below names the reference call site whose arithmetic
 * it replaces (paths relative to the reference checkout, torch/ = installed PyTorch 2.11).
 *
 * Activation layout: NHWC bf16 ("channels-last"); element (b,h,w,c) of a tensor with channel
 * stride cs lives at base + ((b*H + h)*W + w)*cs + c.  Master weights / statistics / sigma: fp32.
 */
#ifndef SGB200_H_
#define SGB200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGB_OK 0
#define SGB_ERR_ARG 1
#define SGB_ERR_CUDA 2
#define SGB_ERR_UNSUPPORTED 3

typedef void* sgb_stream_t; /* cudaStream_t */

/* library / device probe: returns the ABI version (or -1), no GPU work. */
int sgb_abi_version(void);
/* 0 if a sm_100 device is current and usable, else SGB_ERR_CUDA. */
int sgb_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMA-staged NHWC tiles, TMEM accumulators).
 * Stride-1 convolution with KHxKW taps and zero padding; H,W are the OUTPUT spatial size
 * (== input size).  Also serves linear layers (H=W=1) and dgrad (caller passes the
 * transposed/flipped packed weights).
 * Replaces: nn.Conv2d / nn.Linear forward and input-gradient as used by
 *   src/utils/ops.py:165-173,187-188,195-204,219-220 and every conv2d / linear call in
 *   src/models/big_resnet_deep_legacy.py:49-73,210-229, src/models/big_resnet.py, src/models/resnet.py.
 * ------------------------------------------------------------------------------------------ */
typedef struct sgb_conv_desc {
  int32_t B, H, W;
  int32_t Cin, Cout;
  int32_t KH, KW, pad_h, pad_w;
  const void* x;       /* bf16 NHWC input */
  int64_t x_cstride;   /* elements between consecutive pixels (>= Cin, multiple of 8) */
  const void* w;       /* bf16 packed weights [Cout][KH*KW][Cin] */
  float alpha;         /* accumulator scale */
  const float* bias;   /* fp32 [Cout] or NULL */
  const void* residual;/* bf16 NHWC or NULL; added after bias */
  int64_t res_cstride;
  int32_t res_up2;     /* 1: residual is at half resolution, read at (h/2, w/2) (nearest x2 upsample) */
  const void* mask;    /* bf16 NHWC or NULL; output is zeroed where mask <= 0 (ReLU backward) */
  int64_t mask_cstride;
  int32_t relu;        /* 1: ReLU applied last (before mask) */
  void* y;             /* output, bf16 (or fp32 if y_fp32) */
  int64_t y_cstride;
  int32_t y_fp32;
} sgb_conv_desc;

int sgb_conv_fprop(const sgb_conv_desc* d, sgb_stream_t stream);

/* Weight gradient: dw[co][tap][ci] += sum_{b,h,w} dy[b,h,w,co] * x[b,h+kh-pad,w+kw-pad,ci]
 * (tcgen05 with MN-major operands, split over pixels, fp32 atomics into dw).
 * Replaces: conv/linear weight-gradient of torch autograd for the same call sites. */
typedef struct sgb_wgrad_desc {
  int32_t B, H, W;
  int32_t Cin, Cout;
  int32_t KH, KW, pad_h, pad_w;
  const void* x;
  int64_t x_cstride;
  const void* dy;
  int64_t dy_cstride;
  float* dw;          /* fp32 [Cout][KH*KW][Cin] */
  int32_t accumulate; /* 0: dw is zeroed first */
} sgb_wgrad_desc;

int sgb_conv_wgrad(const sgb_wgrad_desc* d, sgb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H_ */
