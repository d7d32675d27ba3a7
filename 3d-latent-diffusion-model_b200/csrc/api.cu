// api.cu - library-wide C ABI plumbing: version, thread-local error text, device check.
#include "common.cuh"

namespace b200ldm {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};
void set_conv_trace(void* p);
void set_conv_trace2(void* p);
void set_attn_trace(void* p);
void set_conv_kernel(int k);
void set_conv_splits(int s);
void set_conv_choice(int c);
void set_conv_trace_sk(void* p);
void set_attn_split(int s);

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
  return d < kMaxDevices ? d : kMaxDevices - 1;
}

int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int d = current_device();
  int v = n[d].load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
    n[d].store(v, std::memory_order_relaxed);
  }
  return v;
}

}  // namespace b200ldm

extern "C" {

int b200ldm_version(void) { return B200LDM_VERSION; }

int b200ldm_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return (int)strlen(b200ldm::t_err);
  strncpy(buf, b200ldm::t_err, n - 1);
  buf[n - 1] = 0;
  return (int)strlen(buf);
}

uint64_t b200ldm_launch_count(void) { return b200ldm::g_launches.load(); }

size_t b200ldm_sizeof_params(int32_t which) {
  return which == 0 ? sizeof(b200ldm_conv3d_params) : which == 1 ? sizeof(b200ldm_sched_params) : 0;
}

int b200ldm_check_device(void) {
  int dev = 0;
  B2_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  B2_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B2_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return b200ldm::fail(B200LDM_UNSUPPORTED_ARCH, "device %d is sm_%d%d; libb200ldm is sm_100a only", dev,
                         major, minor);
  return B200LDM_OK;
}

/* debug only (not in the public header): 8 clock64 stamps per conv CTA are written to `buf` while it is set */
int b200ldm_debug_set_conv_trace(void* buf) {
  b200ldm::set_conv_trace(buf);
  b200ldm::set_conv_trace2(buf);
  b200ldm::set_conv_trace_sk(buf);
  b200ldm::set_attn_trace(buf);
  return B200LDM_OK;
}

/* debug only (not in the public header): pin the conv kernel like B200LDM_CONV_KERNEL does (-1 = measured / model) */
int b200ldm_debug_set_conv_kernel(int k) {
  b200ldm::set_conv_kernel(k);
  return B200LDM_OK;
}

/* debug only (not in the public header): force a split-K factor on the one-CTA-per-tile kernel like B200LDM_CONV_SPLITS
   does (0 = measured / model) */
int b200ldm_debug_set_conv_splits(int s) {
  b200ldm::set_conv_splits(s);
  return B200LDM_OK;
}

/* debug only (not in the public header): force one tuner candidate (choice code, see conv_tc.cu) on every conv; -1 = off */
int b200ldm_debug_set_conv_choice(int c) {
  b200ldm::set_conv_choice(c);
  return B200LDM_OK;
}

/* debug only (not in the public header): attention split-KV factor for the partial wave like B200LDM_ATTN_SPLIT
   (-1 = planner, 0 = never, n >= 2 = force) */
int b200ldm_debug_set_attn_split(int s) {
  b200ldm::set_attn_split(s);
  return B200LDM_OK;
}

int b200ldm_fill_zero(void* p, size_t bytes, b200ldm_stream stream) {
  B2_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
  return B200LDM_OK;
}

}  // extern "C"
