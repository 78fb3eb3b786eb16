// core.cu — status strings, thread-local error detail, DLPack tensor validation.
#include "common.cuh"

namespace od {

static thread_local char g_detail[512] = "";
static unsigned long long g_launches = 0;

static thread_local DeviceScope* g_scope = nullptr;
DeviceScope::DeviceScope() : outer(g_scope) { g_scope = this; }
DeviceScope::~DeviceScope() {
  if (switched) cudaSetDevice(prev);
  g_scope = outer;
}
void DeviceScope::activate(int device_id) {
  DeviceScope* s = g_scope;
  if (!s || s->active) return;
  s->active = true;
  if (cudaGetDevice(&s->prev) != cudaSuccess) return;
  if (s->prev != device_id && cudaSetDevice(device_id) == cudaSuccess) s->switched = true;
}

void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

void set_error_detail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
  va_end(ap);
}

bool is_contiguous(const DLTensor* t) {
  if (!t->strides) return true;
  int64_t expect = 1;
  for (int i = t->ndim - 1; i >= 0; --i) {
    if (t->shape[i] != 1 && t->strides[i] != expect) return false;
    expect *= t->shape[i];
  }
  return true;
}

int check_tensor(const DLTensor* t, const char* name, DType dt, int ndim, bool need_contig, int* device) {
  if (!t) OD_FAIL(OD_ERR_NULL, "%s is NULL", name);
  if (t->device.device_type != kDLCUDA)
    OD_FAIL(OD_ERR_DEVICE, "%s: device_type %d is not kDLCUDA (there is no CPU path)", name, (int)t->device.device_type);
  if (device) {
    if (*device < 0) {
      *device = t->device.device_id;
      DeviceScope::activate(*device);
    } else if (*device != t->device.device_id)
      OD_FAIL(OD_ERR_DEVICE, "%s is on cuda:%d, expected cuda:%d", name, t->device.device_id, *device);
  }
  const uint8_t code = (dt == I32) ? (uint8_t)kDLInt : (uint8_t)kDLFloat;
  const uint8_t bits = (dt == F64) ? 64 : 32;
  if (t->dtype.code != code || t->dtype.bits != bits || t->dtype.lanes != 1)
    OD_FAIL(OD_ERR_DTYPE, "%s: dtype (code %d, bits %d) != expected (code %d, bits %d)", name, t->dtype.code,
            t->dtype.bits, code, bits);
  if (ndim >= 0 && t->ndim != ndim) OD_FAIL(OD_ERR_SHAPE, "%s: rank %d != %d", name, t->ndim, ndim);
  for (int i = 0; i < t->ndim; ++i)
    if (t->shape[i] < 0) OD_FAIL(OD_ERR_SHAPE, "%s: negative extent", name);
  if (need_contig && !is_contiguous(t)) OD_FAIL(OD_ERR_LAYOUT, "%s must be C-contiguous", name);
  if (numel(t) > 0 && dptr<char>(t) == nullptr) OD_FAIL(OD_ERR_NULL, "%s: data pointer is NULL", name);
  return OD_OK;
}

}  // namespace od

extern "C" {

#ifndef OD_SOURCE_HASH
#define OD_SOURCE_HASH "unknown"
#endif
const char* od_source_hash(void) { return OD_SOURCE_HASH; }

int od_version(void) { return 10000 * 0 + 100 * 1 + 0; }

const char* od_strerror(int status) {
  switch (status) {
    case OD_OK: return "ok";
    case OD_ERR_NULL: return "required pointer is NULL";
    case OD_ERR_DTYPE: return "wrong dtype";
    case OD_ERR_SHAPE: return "wrong shape";
    case OD_ERR_DEVICE: return "wrong device (CUDA tensors on one device required)";
    case OD_ERR_LAYOUT: return "tensor must be contiguous / aligned";
    case OD_ERR_WORKSPACE: return "workspace missing or too small";
    case OD_ERR_CUDA: return "CUDA runtime error";
    case OD_ERR_PARAM: return "parameter out of supported range";
    default: return "unknown status";
  }
}

const char* od_last_error_detail(void) { return od::g_detail; }

int64_t od_launch_count(void) { return (int64_t)__atomic_load_n(&od::g_launches, __ATOMIC_RELAXED); }

}  // extern "C"
