// PrismaticImageProcessor.apply_transform up to the uint8 frame (processing_prismatic.py:128-135): optional letterbox
// padding (:23-29), TVF.resize with PIL's bicubic resampling, TVF.center_crop -- for a batch of uint8 HWC frames on the
// device.  (to_tensor + normalize + the bf16 cast follow in preprocess_frames_kernel, elementwise.cu.)
//
// PIL's ImagingResample (Pillow src/libImaging/Resample.c) is integer arithmetic on the pixels: per output index a
// window of the input with bicubic (a = -0.5) weights, evaluated in double precision on the host exactly as
// precompute_coeffs / normalize_coeffs_8bpc do and rounded to 22-bit fixed point; int32 accumulation from 2^21, shift,
// clip to uint8 after the horizontal pass and again after the vertical pass.  The kernels below do the two passes with
// the same tables, so the result is bit-identical to PIL (tests/test_image_transform.py pins it to the reference's own
// processor output).  Only the rows / columns that survive the center crop are computed.
#include <math.h>

#include <vector>

#include "host_util.h"
#include "ops.h"

namespace ovla {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc over the whole axis; tab = per output index [xmin, n, k[0..ksize)]
int resample_table(int in_size, int out_size, std::vector<int>& tab) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  tab.assign(static_cast<size_t>(out_size) * (ksize + 2), 0);
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int* row = &tab[static_cast<size_t>(xx) * (ksize + 2)];
    row[0] = xmin;
    row[1] = xmax;
    for (int x = 0; x < xmax; ++x) {
      double k = w[x];
      if (ww != 0.0) k /= ww;
      row[2 + x] = k < 0 ? static_cast<int>(-0.5 + k * (1 << kPrecisionBits)) : static_cast<int>(0.5 + k * (1 << kPrecisionBits));
    }
  }
  return ksize;
}

__device__ __forceinline__ unsigned char clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[b, r, xo, :] = horizontal pass of padded-source row (row0 + r) at output column (col0 + xo)
__global__ void resize_h_kernel(const unsigned char* __restrict__ src, int H, int W, int pad_x, int pad_y, int fill_rgb,
                                const int* __restrict__ tab_x, int kx, int row0, int n_rows, int col0, int S,
                                unsigned char* __restrict__ tmp, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xo = static_cast<int>(idx % S);
  const int r = static_cast<int>((idx / S) % n_rows);
  const int b = static_cast<int>(idx / (static_cast<long long>(S) * n_rows));
  const int* t = tab_x + static_cast<long long>(col0 + xo) * (kx + 2);
  const int xmin = t[0], n = t[1];
  const int y = row0 + r - pad_y;                 // row in the un-padded frame
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  const bool row_in = y >= 0 && y < H;
  const unsigned char* srow = src + (static_cast<long long>(b) * H + (row_in ? y : 0)) * W * 3;
  const int f0 = fill_rgb & 255, f1 = (fill_rgb >> 8) & 255, f2 = (fill_rgb >> 16) & 255;
  for (int j = 0; j < n; ++j) {
    const int x = xmin + j - pad_x;
    const int k = t[2 + j];
    if (row_in && x >= 0 && x < W) {
      a0 += srow[x * 3] * k;
      a1 += srow[x * 3 + 1] * k;
      a2 += srow[x * 3 + 2] * k;
    } else {
      a0 += f0 * k;
      a1 += f1 * k;
      a2 += f2 * k;
    }
  }
  unsigned char* o = tmp + idx * 3;
  o[0] = clip8(a0);
  o[1] = clip8(a1);
  o[2] = clip8(a2);
}

// out[b, yo, xo, :] = vertical pass over tmp rows at output row (out_row0 + yo)
__global__ void resize_v_kernel(const unsigned char* __restrict__ tmp, int n_rows, int row0, const int* __restrict__ tab_y,
                                int ky, int out_row0, int S, unsigned char* __restrict__ out, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xo = static_cast<int>(idx % S);
  const int yo = static_cast<int>((idx / S) % S);
  const int b = static_cast<int>(idx / (static_cast<long long>(S) * S));
  const int* t = tab_y + static_cast<long long>(out_row0 + yo) * (ky + 2);
  const int ymin = t[0], n = t[1];
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  const unsigned char* p = tmp + ((static_cast<long long>(b) * n_rows + (ymin - row0)) * S + xo) * 3;
  for (int j = 0; j < n; ++j) {
    const int k = t[2 + j];
    a0 += p[0] * k;
    a1 += p[1] * k;
    a2 += p[2] * k;
    p += static_cast<long long>(S) * 3;
  }
  unsigned char* o = out + idx * 3;
  o[0] = clip8(a0);
  o[1] = clip8(a1);
  o[2] = clip8(a2);
}

}  // namespace

// strategy: 0 "resize-naive" (stretch to S x S), 1 "resize-crop" (short side -> S, center crop), 2 "letterbox" (pad to
// square with `fill`, then as resize-crop)
int resize_frames_launch(const void* frames_u8, int B, int H, int W, int strategy, int fill_r, int fill_g, int fill_b,
                         void* out_u8, int S, cudaStream_t st) {
  if (B <= 0) return 0;
  if (H <= 0 || W <= 0 || S <= 0) return set_error("resize_frames: empty frame or output");
  if (strategy < 0 || strategy > 2) return set_error("resize_frames: image resize strategy %d is not supported (0 naive, 1 crop, 2 letterbox)", strategy);
  int pad_x = 0, pad_y = 0;
  if (strategy == 2) {  // processing_prismatic.py:23-29
    const int m = H > W ? H : W;
    pad_x = (m - W) / 2;
    pad_y = (m - H) / 2;
  }
  const int Hp = H + 2 * pad_y, Wp = W + 2 * pad_x;
  int oh = S, ow = S;
  if (strategy != 0) {  // torchvision _compute_resized_output_size: short side -> S, long side -> int(S * long / short)
    if (Wp <= Hp) { ow = S; oh = static_cast<int>(static_cast<long long>(S) * Hp / Wp); }
    else { oh = S; ow = static_cast<int>(static_cast<long long>(S) * Wp / Hp); }
  }
  // torchvision center_crop: int(round((side - S) / 2.0)), Python's round-half-to-even
  auto py_round_half = [](int d) { const int k = d / 2; return (d % 2 == 0) ? k : ((k % 2 == 0) ? k : k + 1); };
  const int top = py_round_half(oh - S), left = py_round_half(ow - S);
  std::vector<int> tx, ty;
  const int kx = resample_table(Wp, ow, tx), ky = resample_table(Hp, oh, ty);
  int row0 = Hp, row1 = 0;  // padded-source rows the cropped output rows read
  for (int yo = top; yo < top + S; ++yo) {
    const int* t = &ty[static_cast<size_t>(yo) * (ky + 2)];
    row0 = t[0] < row0 ? t[0] : row0;
    row1 = t[0] + t[1] > row1 ? t[0] + t[1] : row1;
  }
  const int n_rows = row1 - row0;
  int *d_tx = nullptr, *d_ty = nullptr;
  unsigned char* tmp = nullptr;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_tx), tx.size() * sizeof(int), st));
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_ty), ty.size() * sizeof(int), st));
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&tmp), static_cast<size_t>(B) * n_rows * S * 3, st));
  // pageable sources: these copies return once the data has been staged, so the vectors may die at scope exit
  CUDA_TRY(cudaMemcpyAsync(d_tx, tx.data(), tx.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_ty, ty.data(), ty.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  const int fill = (fill_r & 255) | ((fill_g & 255) << 8) | ((fill_b & 255) << 16);
  const long long n1 = static_cast<long long>(B) * n_rows * S, n2 = static_cast<long long>(B) * S * S;
  ProfScope prof(kCatOther, 0.0, 3.0 * B * (static_cast<double>(H) * W + 2.0 * n_rows * S + static_cast<double>(S) * S), st);
  resize_h_kernel<<<static_cast<unsigned>((n1 + 255) / 256), 256, 0, st>>>(static_cast<const unsigned char*>(frames_u8), H, W,
                                                                            pad_x, pad_y, fill, d_tx, kx, row0, n_rows, left, S,
                                                                            tmp, n1);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  resize_v_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, st>>>(tmp, n_rows, row0, d_ty, ky, top, S,
                                                                            static_cast<unsigned char*>(out_u8), n2);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  CUDA_TRY(cudaFreeAsync(tmp, st));
  CUDA_TRY(cudaFreeAsync(d_ty, st));
  CUDA_TRY(cudaFreeAsync(d_tx, st));
  return 0;
}

}  // namespace ovla
