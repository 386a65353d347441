// C-ABI dispatch: picks the tcgen05 kernel when the shape tiles (channels % 64, 128-pixel boxes) and the
// CUDA-core kernel otherwise.  There is no CPU path anywhere in this library.
#include "common.cuh"

#include <stdlib.h>

namespace bd {
void count_launch(int n);
int simt_conv_launch(const bd_conv_args* a, bool dgrad, cudaStream_t st);
int simt_wgrad_launch(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, float* dw, float* dbias, int B, int H,
                      int W, int Cin, int Cout, int ksize, int mode, int pad_in, int accumulate, cudaStream_t st);
int attn_fwd_simt(const void* qkv, int64_t ld_qkv, void* probs, void* out, int64_t ld_out, int B, int S, int C,
                  int heads, float scale, cudaStream_t st);
int attn_bwd_simt(const void* qkv, int64_t ld_qkv, const void* probs, const void* d_out, int64_t ld_dout, void* d_qkv,
                  int64_t ld_dqkv, void* work, int B, int S, int C, int heads, float scale, cudaStream_t st);
int softmax_fwd_launch(const float* scores, void* probs, int64_t rows, int S, float scale, cudaStream_t st);
int softmax_bwd_launch(const float* dP, const void* P, void* dS, int64_t rows, int S, float scale, cudaStream_t st);
__global__ void zero_f32_kernel(float* p, size_t n);

namespace umma {
struct FpropCall {
  const void* a; int64_t ld_a; int Ca;
  const void* a2; int64_t ld_a2; int Ca2;
  int NB, H, W;
  const void* b; int b_rows, b_cols, b_z; int64_t ld_b;
  const void* b2; int64_t ld_b2;
  int N, ks; bool b_mn; bool batched; bool flip_taps;
  const float* bias; const float* bias2; const float* rowbias; int64_t ld_rowbias; int HW_rowbias;
  const void* residual; int64_t ld_res; float scale; void* y; int64_t ld_y; int out_f32;
  int a_stride;
  int a_H, a_W;
  int ntap_override;
  int tap_dx[9], tap_dy[9], tap_z[9];
  int64_t out_sn, out_sh, out_sw, out_off;
  float* gn_sums; int64_t ld_sums;
};
struct WgradCall {
  const void* a; int64_t ld_a; int Mtot;
  const void* b; int64_t ld_b; int Ntot;
  int NB, H, W, ks; bool batched;
  void* y; int64_t ld_y; int out_mode;
  int zcount;
  int b_stride, b_pad, b_H, b_W;
};
struct Conv3Call {
  const void* a; int64_t ld_a; int Ca;
  const void* a2; int64_t ld_a2; int Ca2;
  int NB, H, W;
  const void* b; int64_t ld_b; int b_rows;
  const void* b2; int64_t ld_b2;
  int N; bool b_mn; bool flip;
  const float* bias; const float* bias2; const float* rowbias; int64_t ld_rowbias;
  const void* residual; int64_t ld_res; float scale; void* y; int64_t ld_y; int out_f32;
  float* gn_sums; int64_t ld_sums;
};
struct Wgrad3Call {
  const void* dy; int64_t ld_dy; int Cout;
  const void* x; int64_t ld_x; int Cin;
  int NB, H, W;
  float* dw;
};
int wgrad3_supported(const Wgrad3Call& c);
int wgrad3_launch(const Wgrad3Call& c, cudaStream_t st);
int conv3_supported(const Conv3Call& c);
int conv3_gn_sums_supported(const Conv3Call& c);
int conv3_launch(const Conv3Call& c, cudaStream_t st);
int fprop_supported(const FpropCall& c);
int fprop_gn_sums_supported(const FpropCall& c);
int fprop_launch(const FpropCall& c, cudaStream_t st);
int wgrad_supported(const WgradCall& c);
int wgrad_launch(const WgradCall& c, cudaStream_t st);
int read_error_flag();
int* error_flag();
bool attn_fused_supported(int S, int C, int heads, int64_t ld_a, int64_t ld_qkv);
int attn_fused_launch(int mode, const void* a, int64_t ld_a, const void* b1, const void* b2, int64_t ld_qkv,
                      const void* probs_in, void* p_out, void* y, int64_t ld_y, int B, int S, int C, float scale,
                      cudaStream_t st);
int attn_dkv_launch(const void* probs, const void* ds, const void* d_out, int64_t ld_dout, const void* q, int64_t ld_qkv,
                    void* dk, void* dv, int64_t ld_y, int B, int S, int C, cudaStream_t st);
}  // namespace umma

static bool umma_allowed() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BD_FORCE_SIMT");
    v = (e && e[0] == '1') ? 0 : (bd_device_supported() ? 1 : 0);
  }
  return v == 1;
}

__global__ void __launch_bounds__(256) cast_f32_f16_kernel(const float* __restrict__ s, __half* __restrict__ d, size_t n) {
  size_t n8 = n / 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(s)[2 * i], b = reinterpret_cast<const float4*>(s)[2 * i + 1];
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    reinterpret_cast<half8*>(d)[i] = pack8(f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) d[n8 * 8 + threadIdx.x] = __float2half_rn(s[n8 * 8 + threadIdx.x]);
}
__global__ void __launch_bounds__(256) silu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           float* __restrict__ dx, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float z = x[i], sg = sigmoid_f(z);
    dx[i] = dy[i] * sg * (1.0f + z * (1.0f - sg));
  }
}
__global__ void __launch_bounds__(256) silu_f32_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(silu_f(x[i]));
}
__global__ void __launch_bounds__(256) pack_weight2_kernel(const float* __restrict__ w, float* __restrict__ wf32,
                                                           __half* __restrict__ wf16, int O, int I, int taps) {
  size_t n = (size_t)O * I * taps;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int tap = (int)(i % taps);
    size_t r = i / taps;
    int ci = (int)(r % I), o = (int)(r / I);
    size_t dst = ((size_t)tap * O + o) * I + ci;
    if (wf32) wf32[dst] = w[i];
    if (wf16) wf16[dst] = __float2half_rn(w[i]);
  }
}

static int conv_out_hw(const bd_conv_args* a, int* Ho, int* Wo) {
  if (a->mode == BD_CONV_S2_PAD01) {
    *Ho = (a->H + (a->pad ? 2 : 1) - 3) / 2 + 1;
    *Wo = (a->W + (a->pad ? 2 : 1) - 3) / 2 + 1;
  } else {
    *Ho = a->H;
    *Wo = a->W;
  }
  return 0;
}

static int check_conv(const bd_conv_args* a, const char* who) {
  BD_CHECK_ARG(a && a->x && a->w && a->y, "%s: null pointer", who);
  BD_CHECK_ARG(a->B >= 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "%s: bad shape", who);
  BD_CHECK_ARG(a->ksize == 1 || a->ksize == 3, "%s: ksize must be 1 or 3", who);
  BD_CHECK_ARG(a->mode == BD_CONV_S1 || (a->mode == BD_CONV_S2_PAD01 && a->ksize == 3), "%s: bad mode", who);
  BD_CHECK_ARG(a->Cin % 8 == 0 && a->Cout % 8 == 0 && a->ld_x % 8 == 0 && a->ld_y % 4 == 0,
               "%s: channels / ld must be multiples of 8 (Cin=%d Cout=%d)", who, a->Cin, a->Cout);
  BD_CHECK_ARG(!a->x2 || (a->w2 && a->Cin2 % 8 == 0 && a->ld_x2 % 8 == 0), "%s: bad second segment", who);
  return BD_OK;
}

}  // namespace bd

namespace bd { namespace umma { const __half* conv3_identity(); } }
using namespace bd;

extern "C" {

int bd_init(void) {
  int* f = umma::error_flag();
  if (!f) { set_error("bd_init: cudaMalloc failed"); return BD_ERR_CUDA; }
  (void)num_sms();
  (void)umma::conv3_identity();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bd_init: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
  return BD_OK;
}
int bd_umma_error(void) { return umma::read_error_flag(); }

static umma::FpropCall fprop_call_fwd(const bd_conv_args* a) {
  umma::FpropCall c;
  memset(&c, 0, sizeof(c));
  c.a = a->x; c.ld_a = a->ld_x; c.Ca = a->Cin;
  c.a2 = a->x2; c.ld_a2 = a->ld_x2; c.Ca2 = a->Cin2;
  c.NB = a->B; c.H = a->H; c.W = a->W;
  c.b = a->w; c.b_rows = a->Cout; c.b_cols = a->Cin; c.b_z = a->ksize * a->ksize; c.ld_b = a->Cin;
  c.b2 = a->w2; c.ld_b2 = a->Cin2;
  c.N = a->Cout; c.ks = a->ksize; c.b_mn = false; c.batched = false; c.flip_taps = false;
  c.bias = a->bias; c.bias2 = a->bias2; c.rowbias = a->rowbias; c.ld_rowbias = a->ld_rowbias; c.HW_rowbias = a->H * a->W;
  c.residual = a->residual; c.ld_res = a->ld_res; c.scale = a->out_scale; c.y = a->y; c.ld_y = a->ld_y;
  c.out_f32 = a->out_dtype == BD_OUT_F32;
  if (a->mode == BD_CONV_S2_PAD01) {
    // Downsample2D: the A tile is sampled at every other pixel by the TMA unit (elementStrides = 2); tiles run over
    // the OUTPUT grid.  tap (r, s) reads input (2*ho + r - pad, 2*wo + s - pad); out-of-bounds -> zero fill.
    int Ho, Wo;
    conv_out_hw(a, &Ho, &Wo);
    c.a_stride = 2; c.a_H = a->H; c.a_W = a->W; c.H = Ho; c.W = Wo; c.HW_rowbias = Ho * Wo;
    c.ntap_override = 9;
    for (int t = 0; t < 9; ++t) { c.tap_dy[t] = t / 3 - a->pad; c.tap_dx[t] = t % 3 - a->pad; c.tap_z[t] = t; }
  }
  return c;
}

static umma::Conv3Call conv3_call_fwd(const bd_conv_args* a) {
  umma::Conv3Call h;
  memset(&h, 0, sizeof(h));
  h.a = a->x; h.ld_a = a->ld_x; h.Ca = a->Cin; h.a2 = a->x2; h.ld_a2 = a->ld_x2; h.Ca2 = a->Cin2;
  h.NB = a->B; h.H = a->H; h.W = a->W; h.b = a->w; h.ld_b = a->Cin; h.b_rows = a->Cout; h.b2 = a->w2; h.ld_b2 = a->Cin2;
  h.N = a->Cout; h.b_mn = false; h.flip = false;
  h.bias = a->bias; h.bias2 = a->bias2; h.rowbias = a->rowbias; h.ld_rowbias = a->ld_rowbias;
  h.residual = a->residual; h.ld_res = a->ld_res; h.scale = a->out_scale; h.y = a->y; h.ld_y = a->ld_y;
  h.out_f32 = a->out_dtype == BD_OUT_F32;
  h.gn_sums = a->gn_sums; h.ld_sums = a->ld_sums;
  return h;
}

int bd_conv_fwd(const bd_conv_args* a, void* stream) {
  int rc = check_conv(a, "bd_conv_fwd");
  if (rc) return rc;
  if (a->B == 0) return BD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  umma::FpropCall c = fprop_call_fwd(a);
  const bool s2ok = a->mode == BD_CONV_S1 || (a->H % 2 == 0 && a->W % 2 == 0 && !a->x2);
  const bool can = s2ok && a->ld_y % 8 == 0 && (!a->residual || a->ld_res % 8 == 0) && umma::fprop_supported(c);
  if ((a->impl == BD_IMPL_UMMA || a->impl == BD_IMPL_UMMA_TILE) && !(can && bd_device_supported())) {
    set_error("bd_conv_fwd: tcgen05 path requested but shape unsupported (Cin=%d Cout=%d H=%d W=%d mode=%d)", a->Cin, a->Cout, a->H, a->W, a->mode);
    return BD_ERR_UNSUPPORTED;
  }
  if (a->impl == BD_IMPL_UMMA || a->impl == BD_IMPL_UMMA_TILE || (a->impl == BD_IMPL_AUTO && can && umma_allowed())) {
    umma::Conv3Call h = conv3_call_fwd(a);
    if (a->impl != BD_IMPL_UMMA_TILE && a->mode == BD_CONV_S1 && a->ksize == 3 && umma::conv3_supported(h)) {
      rc = umma::conv3_launch(h, st);   // halo-reuse kernel (umma_conv3.cu)
    } else {
      c.gn_sums = a->gn_sums; c.ld_sums = a->ld_sums;   // generic kernels: statistics in epilogue_warp (rejected there if unsupported)
      rc = umma::fprop_launch(c, st);
    }
    if (rc) return rc;
  } else {
    if (a->gn_sums) { set_error("bd_conv_fwd: gn_sums is not supported on the CUDA-core path"); return BD_ERR_UNSUPPORTED; }
    simt_conv_launch(a, false, st);
  }
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_conv_fwd_gn_sums_supported(const bd_conv_args* a) {
  if (!a || a->impl == BD_IMPL_SIMT || check_conv(a, "bd_conv_fwd_gn_sums_supported")) return 0;
  if (a->out_dtype != BD_OUT_F16 || !umma_allowed()) return 0;
  umma::Conv3Call h = conv3_call_fwd(a);
  const bool halo = a->impl != BD_IMPL_UMMA_TILE && a->mode == BD_CONV_S1 && a->ksize == 3 && umma::conv3_supported(h);
  if (halo) return umma::conv3_gn_sums_supported(h);   // bd_conv_fwd takes the halo-reuse kernels for this call
  // otherwise the generic tcgen05 kernels, if bd_conv_fwd would run them at all (same conditions as there)
  umma::FpropCall c = fprop_call_fwd(a);
  const bool s2ok = a->mode == BD_CONV_S1 || (a->H % 2 == 0 && a->W % 2 == 0 && !a->x2);
  const bool can = s2ok && a->ld_y % 8 == 0 && (!a->residual || a->ld_res % 8 == 0) && umma::fprop_supported(c);
  return can && umma::fprop_gn_sums_supported(c);
}

int bd_conv_dgrad(const bd_conv_args* a, void* stream) {
  int rc = check_conv(a, "bd_conv_dgrad");
  if (rc) return rc;
  if (a->B == 0) return BD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo;
  conv_out_hw(a, &Ho, &Wo);
  umma::FpropCall c;
  memset(&c, 0, sizeof(c));
  c.a = a->x; c.ld_a = a->ld_x; c.Ca = a->Cout;   // x := dy (B,Ho,Wo,Cout)
  c.NB = a->B; c.H = a->H; c.W = a->W;
  c.b = a->w; c.b_rows = a->Cout; c.b_cols = a->Cin; c.b_z = a->ksize * a->ksize; c.ld_b = a->Cin;
  c.N = a->Cin; c.ks = a->ksize; c.b_mn = true; c.batched = false; c.flip_taps = true;
  c.bias = nullptr; c.residual = a->residual; c.ld_res = a->ld_res; c.scale = a->out_scale; c.y = a->y; c.ld_y = a->ld_y;
  c.out_f32 = a->out_dtype == BD_OUT_F32;
  const bool s2 = a->mode == BD_CONV_S2_PAD01;
  if (s2) { c.H = Ho; c.W = Wo; }  // tiles run over the dY grid; each launch fills one parity class of dX
  const bool s2ok = !s2 || (a->H == 2 * Ho && a->W == 2 * Wo);
  const bool can = s2ok && a->ld_y % 8 == 0 && (!a->residual || a->ld_res % 8 == 0) && umma::fprop_supported(c);
  if ((a->impl == BD_IMPL_UMMA || a->impl == BD_IMPL_UMMA_TILE) && !(can && bd_device_supported())) {
    set_error("bd_conv_dgrad: tcgen05 path requested but shape unsupported");
    return BD_ERR_UNSUPPORTED;
  }
  if (a->impl == BD_IMPL_UMMA || a->impl == BD_IMPL_UMMA_TILE || (a->impl == BD_IMPL_AUTO && can && umma_allowed())) {
    if (!s2) {
      umma::Conv3Call h;
      memset(&h, 0, sizeof(h));
      h.a = a->x; h.ld_a = a->ld_x; h.Ca = a->Cout; h.NB = a->B; h.H = a->H; h.W = a->W;
      h.b = a->w; h.ld_b = a->Cin; h.b_rows = a->Cout; h.N = a->Cin; h.b_mn = true; h.flip = true;
      h.residual = a->residual; h.ld_res = a->ld_res; h.scale = a->out_scale; h.y = a->y; h.ld_y = a->ld_y; h.out_f32 = c.out_f32;
      if (a->impl != BD_IMPL_UMMA_TILE && a->ksize == 3 && umma::conv3_supported(h))
        rc = umma::conv3_launch(h, st);
      else
        rc = umma::fprop_launch(c, st);
      if (rc) return rc;
    } else {
      // dX[2i+ah, 2j+aw] = sum over taps (r, s) with (ah + pad - r), (aw + pad - s) even of
      //                    dY[i + (ah + pad - r)/2, j + (aw + pad - s)/2] W[r, s]^T
      for (int ah = 0; ah < 2; ++ah)
        for (int aw = 0; aw < 2; ++aw) {
          umma::FpropCall q = c;
          q.ntap_override = 0;
          for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
              const int nh = ah + a->pad - r, nw = aw + a->pad - s;
              if ((nh & 1) || (nw & 1)) continue;
              const int t = q.ntap_override++;
              q.tap_dy[t] = nh / 2; q.tap_dx[t] = nw / 2; q.tap_z[t] = r * 3 + s;
            }
          q.out_sn = (int64_t)a->H * a->W; q.out_sh = 2 * (int64_t)a->W; q.out_sw = 2;
          q.out_off = (int64_t)ah * a->W + aw;
          rc = umma::fprop_launch(q, st);
          if (rc) return rc;
        }
    }
  } else {
    simt_conv_launch(a, true, st);
  }
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_conv_wgrad(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, float* dw, float* dbias, int B, int H,
                  int W, int Cin, int Cout, int ksize, int mode, int pad, int accumulate, int impl, void* stream) {
  BD_CHECK_ARG(x && dy && dw, "bd_conv_wgrad: null pointer");
  BD_CHECK_ARG((ksize == 1 || ksize == 3) && Cin % 8 == 0 && Cout % 8 == 0 && ld_x % 8 == 0 && ld_dy % 8 == 0,
               "bd_conv_wgrad: bad shape");
  if (B == 0) return BD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  umma::WgradCall c;
  memset(&c, 0, sizeof(c));
  c.a = dy; c.ld_a = ld_dy; c.Mtot = Cout;
  c.b = x; c.ld_b = ld_x; c.Ntot = Cin;
  c.NB = B; c.H = H; c.W = W; c.ks = ksize; c.batched = false;
  c.y = dw; c.ld_y = Cin; c.out_mode = 0; c.zcount = ksize * ksize;
  int64_t npix = (int64_t)B * H * W;
  bool s2ok = true;
  if (mode == BD_CONV_S2_PAD01) {
    // k-blocks run over the dY grid; X is sampled with stride 2 by the TMA unit
    const int Ho = (H + (pad ? 2 : 1) - 3) / 2 + 1, Wo = (W + (pad ? 2 : 1) - 3) / 2 + 1;
    s2ok = ksize == 3 && H == 2 * Ho && W == 2 * Wo;
    c.b_stride = 2; c.b_pad = pad; c.b_H = H; c.b_W = W; c.H = Ho; c.W = Wo;
    npix = (int64_t)B * Ho * Wo;
  }
  const bool can = s2ok && umma::wgrad_supported(c);
  if ((impl == BD_IMPL_UMMA || impl == BD_IMPL_UMMA_TILE) && !(can && bd_device_supported())) {
    set_error("bd_conv_wgrad: tcgen05 path requested but shape unsupported");
    return BD_ERR_UNSUPPORTED;
  }
  if (impl == BD_IMPL_UMMA || impl == BD_IMPL_UMMA_TILE || (impl == BD_IMPL_AUTO && can && umma_allowed())) {
    const size_t n = (size_t)Cout * Cin * ksize * ksize;
    if (!accumulate) {
      zero_f32_kernel<<<ceil_div(n, 2048), 256, 0, st>>>(dw, n);
      count_launch(1);
    }
    umma::Wgrad3Call w3{dy, ld_dy, Cout, x, ld_x, Cin, B, H, W, dw};
    int rc;
    if (impl != BD_IMPL_UMMA_TILE && mode == BD_CONV_S1 && ksize == 3 && umma::wgrad3_supported(w3))
      rc = umma::wgrad3_launch(w3, st);   // 3 taps per CTA sharing one dY / X-halo tile (umma_wgrad3.cu)
    else
      rc = umma::wgrad_launch(c, st);
    if (rc) return rc;
    if (dbias) {
      rc = bd_colsum_f16(dy, ld_dy, dbias, 0, 1, npix, Cout, accumulate, stream);
      if (rc) return rc;
    }
  } else {
    simt_wgrad_launch(x, ld_x, dy, ld_dy, dw, dbias, B, H, W, Cin, Cout, ksize, mode, pad, accumulate, st);
  }
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_pack_conv_weight(const float* w_oihw, float* w_packed_f32, void* w_packed_f16, int O, int I, int ksize,
                        void* stream) {
  BD_CHECK_ARG(w_oihw && (w_packed_f32 || w_packed_f16) && O > 0 && I > 0 && (ksize == 1 || ksize == 3), "bd_pack_conv_weight: bad argument");
  size_t n = (size_t)O * I * ksize * ksize;
  pack_weight2_kernel<<<ceil_div(n, 1024), 256, 0, (cudaStream_t)stream>>>(w_oihw, w_packed_f32, (__half*)w_packed_f16, O, I, ksize * ksize);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}
int bd_cast_f32_to_f16(const float* src, void* dst, size_t n, void* stream) {
  BD_CHECK_ARG(src && dst && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "bd_cast_f32_to_f16: pointers must be 16-byte aligned");
  if (n == 0) return BD_OK;
  int grid = ceil_div(n / 8 + 1, 256);
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  cast_f32_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst, n);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}
int bd_silu_bwd_f32(const float* dy, const float* x, float* dx, size_t n, void* stream) {
  BD_CHECK_ARG(dy && x && dx, "bd_silu_bwd_f32: null pointer");
  if (n == 0) return BD_OK;
  silu_bwd_f32_kernel<<<ceil_div(n, 256) < 1024 ? ceil_div(n, 256) : 1024, 256, 0, (cudaStream_t)stream>>>(dy, x, dx, n);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}
int bd_silu_f32_to_f16(const float* x, void* y, size_t n, void* stream) {
  BD_CHECK_ARG(x && y, "bd_silu_f32_to_f16: null pointer");
  if (n == 0) return BD_OK;
  silu_f32_f16_kernel<<<ceil_div(n, 256) < 1024 ? ceil_div(n, 256) : 1024, 256, 0, (cudaStream_t)stream>>>(x, (__half*)y, n);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

// ---------------------------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------------------------
static bool attn_umma_ok(int B, int S, int C, int heads, int64_t ld_qkv) {
  return heads == 1 && S % 128 == 0 && C % 64 == 0 && ld_qkv % 8 == 0 && S <= 1024;
}
size_t bd_attention_fwd_workspace_bytes(int B, int S, int C, int heads) {
  return (size_t)B * heads * S * S * sizeof(float);
}
size_t bd_attention_bwd_workspace_bytes(int B, int S, int C, int heads) {
  return (size_t)B * heads * S * S * (sizeof(float) + sizeof(__half)) + 256;
}

int bd_attention_fwd(const void* qkv, int64_t ld_qkv, void* probs, void* out, int64_t ld_out, void* work, int B, int S,
                     int C, int heads, float scale, int impl, void* stream) {
  BD_CHECK_ARG(qkv && out && heads > 0 && C % heads == 0 && (C / heads) % 8 == 0 && ld_qkv % 8 == 0 && ld_out % 8 == 0,
               "bd_attention_fwd: bad argument (C=%d heads=%d)", C, heads);
  if (B == 0) return BD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool can = attn_umma_ok(B, S, C, heads, ld_qkv) && work && probs;
  if (impl == BD_IMPL_UMMA && !(can && bd_device_supported())) {
    set_error("bd_attention_fwd: tcgen05 path needs heads==1, S %% 128 == 0, C %% 64 == 0, work and probs buffers");
    return BD_ERR_UNSUPPORTED;
  }
  if ((impl == BD_IMPL_UMMA || (impl == BD_IMPL_AUTO && can && umma_allowed())) &&
      umma::attn_fused_supported(S, C, heads, ld_qkv, ld_qkv)) {
    // one kernel: Q K^T -> softmax -> P V (umma_attn.cu); probs (fp16) is still written for the backward
    const __half* q = (const __half*)qkv;
    int rc = umma::attn_fused_launch(0, q, ld_qkv, q + C, q + 2 * C, ld_qkv, nullptr, probs, out, ld_out, B, S, C, scale, st);
    if (rc) return rc;
  } else if (impl == BD_IMPL_UMMA || (impl == BD_IMPL_AUTO && can && umma_allowed())) {
    const __half* q = (const __half*)qkv;
    umma::FpropCall c;
    memset(&c, 0, sizeof(c));
    // scores = Q K^T  (fp32)
    c.a = q; c.ld_a = ld_qkv; c.Ca = C; c.NB = B; c.H = 1; c.W = S;
    c.b = q + C; c.b_rows = S; c.b_cols = C; c.b_z = B; c.ld_b = ld_qkv;
    c.N = S; c.ks = 1; c.b_mn = false; c.batched = true; c.scale = 1.0f; c.y = work; c.ld_y = S; c.out_f32 = 1;
    int rc = umma::fprop_launch(c, st);
    if (rc) return rc;
    softmax_fwd_launch((const float*)work, probs, (int64_t)B * S, S, scale, st);
    // out = P V   (V is an MN-major operand, read in place)
    memset(&c, 0, sizeof(c));
    c.a = probs; c.ld_a = S; c.Ca = S; c.NB = B; c.H = 1; c.W = S;
    c.b = q + 2 * C; c.b_rows = S; c.b_cols = C; c.b_z = B; c.ld_b = ld_qkv;
    c.N = C; c.ks = 1; c.b_mn = true; c.batched = true; c.scale = 1.0f; c.y = out; c.ld_y = ld_out; c.out_f32 = 0;
    rc = umma::fprop_launch(c, st);
    if (rc) return rc;
  } else {
    if (attn_fwd_simt(qkv, ld_qkv, probs, out, ld_out, B, S, C, heads, scale, st)) {
      set_error("bd_attention_fwd: sequence / head_dim too large for the CUDA-core kernel (S=%d d=%d)", S, C / heads);
      return BD_ERR_UNSUPPORTED;
    }
  }
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_attention_bwd(const void* qkv, int64_t ld_qkv, const void* probs, const void* d_out, int64_t ld_dout,
                     void* d_qkv, int64_t ld_dqkv, void* work, int B, int S, int C, int heads, float scale, int impl,
                     void* stream) {
  BD_CHECK_ARG(qkv && probs && d_out && d_qkv && work && heads > 0 && C % heads == 0 && (C / heads) % 8 == 0,
               "bd_attention_bwd: bad argument");
  BD_CHECK_ARG(ld_qkv % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "bd_attention_bwd: ld must be a multiple of 8");
  if (B == 0) return BD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool can = attn_umma_ok(B, S, C, heads, ld_qkv);
  if (impl == BD_IMPL_UMMA && !(can && bd_device_supported())) {
    set_error("bd_attention_bwd: tcgen05 path needs heads==1, S %% 128 == 0, C %% 64 == 0");
    return BD_ERR_UNSUPPORTED;
  }
  if (impl == BD_IMPL_UMMA || (impl == BD_IMPL_AUTO && can && umma_allowed())) {
    const __half* q = (const __half*)qkv;
    __half* dq = (__half*)d_qkv;
    float* dP = (float*)work;
    __half* dS = (__half*)((char*)work + (((size_t)B * S * S * sizeof(float) + 255) & ~(size_t)255));
    int rc;
    if (umma::attn_fused_supported(S, C, heads, ld_dout, ld_qkv)) {
      // one kernel: dP = dO V^T -> dS = scale P (dP - rowsum(dP P)) -> dQ = dS K; dS (fp16) also goes to `work` for dK
      rc = umma::attn_fused_launch(1, d_out, ld_dout, q + 2 * C, q + C, ld_qkv, probs, dS, dq, ld_dqkv, B, S, C, scale, st);
      if (rc) return rc;
    } else {
      umma::FpropCall c;
      memset(&c, 0, sizeof(c));
      // dP = dO V^T
      c.a = d_out; c.ld_a = ld_dout; c.Ca = C; c.NB = B; c.H = 1; c.W = S;
      c.b = q + 2 * C; c.b_rows = S; c.b_cols = C; c.b_z = B; c.ld_b = ld_qkv;
      c.N = S; c.ks = 1; c.b_mn = false; c.batched = true; c.scale = 1.0f; c.y = dP; c.ld_y = S; c.out_f32 = 1;
      rc = umma::fprop_launch(c, st);
      if (rc) return rc;
      softmax_bwd_launch(dP, probs, dS, (int64_t)B * S, S, scale, st);
      // dQ = dS K
      memset(&c, 0, sizeof(c));
      c.a = dS; c.ld_a = S; c.Ca = S; c.NB = B; c.H = 1; c.W = S;
      c.b = q + C; c.b_rows = S; c.b_cols = C; c.b_z = B; c.ld_b = ld_qkv;
      c.N = C; c.ks = 1; c.b_mn = true; c.batched = true; c.scale = 1.0f; c.y = dq; c.ld_y = ld_dqkv; c.out_f32 = 0;
      rc = umma::fprop_launch(c, st);
      if (rc) return rc;
    }
    if (umma::attn_fused_supported(S, C, heads, ld_dout, ld_qkv) && !getenv("BD_NO_ATTN_DKV")) {
      // dK = dS^T Q and dV = P^T dO: one launch, both accumulators in TMEM (umma_attn.cu)
      rc = umma::attn_dkv_launch(probs, dS, d_out, ld_dout, q, ld_qkv, dq + C, dq + 2 * C, ld_dqkv, B, S, C, st);
      if (rc) return rc;
      BD_CHECK_LAUNCH();
      return BD_OK;
    }
    // dK = dS^T Q ; dV = P^T dO
    umma::WgradCall w;
    memset(&w, 0, sizeof(w));
    w.a = dS; w.ld_a = S; w.Mtot = S; w.b = q; w.ld_b = ld_qkv; w.Ntot = C;
    w.NB = B; w.H = 1; w.W = S; w.ks = 1; w.batched = true; w.y = dq + C; w.ld_y = ld_dqkv; w.out_mode = 2; w.zcount = B;
    rc = umma::wgrad_launch(w, st);
    if (rc) return rc;
    w.a = probs; w.b = d_out; w.ld_b = ld_dout; w.y = dq + 2 * C;
    rc = umma::wgrad_launch(w, st);
    if (rc) return rc;
  } else {
    if (attn_bwd_simt(qkv, ld_qkv, probs, d_out, ld_dout, d_qkv, ld_dqkv, work, B, S, C, heads, scale, st)) {
      set_error("bd_attention_bwd: sequence / head_dim too large for the CUDA-core kernel");
      return BD_ERR_UNSUPPORTED;
    }
  }
  BD_CHECK_LAUNCH();
  return BD_OK;
}

}  // extern "C"
