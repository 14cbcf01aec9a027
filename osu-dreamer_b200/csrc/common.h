// Host-side common pieces: error reporting, TMA tensor-map encoding (driver entry point fetched at
// run time so the library links without libcuda on the GPU-less build box), launch helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace osd {

// thread-local error string returned by osd_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define OSD_CHECK(cond, ...)      \
  do {                            \
    if (!(cond)) {                \
      osd::set_error(__VA_ARGS__); \
      return 1;                   \
    }                             \
  } while (0)

#define OSD_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      osd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                              \
    }                                                                                        \
  } while (0)

// after every kernel launch: count it (osd_launch_count) and surface launch-configuration errors
#define OSD_LAUNCHED()                   \
  do {                                   \
    osd::count_launch();                 \
    OSD_CUDA(cudaGetLastError());        \
  } while (0)

#define OSD_TRY(expr)      \
  do {                     \
    int _s = (expr);       \
    if (_s != 0) return _s; \
  } while (0)

int num_sms();
// function attributes (cudaFuncSetAttribute) are per device: a `static DeviceOnce once; if (once.first()) {...}` guard
// runs its body once for every device the process touches
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
void count_launch();
unsigned long long launch_count();

// 2D/3D tiled tensor map, SWIZZLE_128B, element size 2 (bf16) or 4 (fp32/tf32).
// dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of dims[i+1].
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box);

inline int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {row_stride_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return make_tmap(out, base, elem_bytes, 2, dims, strides, box);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace osd
