// srj_jni_common.hpp -- what the two JNI translation units share: status -> Java exception mapping (the classes
// the reference throws, src/main/cpp/src/error.hpp:181-239), a per-schema plan cache, column_view -> srj_column.
#pragma once
#ifdef SRJ_JNI_STUBS
#include "jni_stub.h"
#include "cudf_abi_stub.hpp"
#else
#include <jni.h>
#include <cudf/column/column.hpp>
#include <cudf/column/column_factories.hpp>
#include <cudf/column/column_view.hpp>
#include <cudf/table/table_view.hpp>
#include <cudf/utilities/default_stream.hpp>
#include <rmm/device_buffer.hpp>
#include "cudf_jni_apis.hpp"
#endif

#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/srj_b200.h"

namespace srjshim {

// SRJ_EINVAL / SRJ_EUNSUPPORTED -> CudfException (cudf::logic_error), SRJ_EOVERFLOW -> CudfColumnSizeOverflowException,
// SRJ_ENOMEM -> OutOfMemoryError, SRJ_ECUDA -> CudaException (error.hpp:181-239).  Returns true when it threw.
inline bool throw_if_error(JNIEnv* env, int status)
{
  if (status == SRJ_OK) return false;
  const char* cls = "ai/rapids/cudf/CudfException";
  if (status == SRJ_EOVERFLOW) cls = "ai/rapids/cudf/CudfColumnSizeOverflowException";
  else if (status == SRJ_ENOMEM) cls = "java/lang/OutOfMemoryError";
  else if (status == SRJ_ECUDA) cls = "ai/rapids/cudf/CudaException";
  if (!env->ExceptionCheck()) env->ThrowNew(env->FindClass(cls), srj_last_error());
  return true;
}

inline void throw_java(JNIEnv* env, const char* cls, const char* msg)
{
  if (!env->ExceptionCheck()) env->ThrowNew(env->FindClass(cls), msg);
}

// One srj_plan per (device, schema): plans are immutable and thread-safe, so concurrent Spark tasks share them.
inline const srj_plan* plan_for(const std::vector<int32_t>& types, const std::vector<int32_t>& scales, int* status)
{
  static std::mutex mu;
  static std::map<std::pair<std::vector<int32_t>, std::vector<int32_t>>, srj_plan*> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_pair(types, scales);
  auto it  = cache.find(key);
  if (it != cache.end()) { *status = SRJ_OK; return it->second; }
  srj_plan* p = nullptr;
  *status     = srj_plan_create(types.data(), scales.data(), static_cast<int32_t>(types.size()), &p);
  if (*status == SRJ_OK) cache.emplace(std::move(key), p);
  return p;
}

// The cudf::column_view fields the C ABI reads (include/srj_b200.h: srj_column).  Sliced views are not supported,
// as in the reference (RC:1809-1811 passes raw null_mask()).
inline srj_column to_srj(const cudf::column_view& c)
{
  srj_column s{};
  s.type_id   = static_cast<int32_t>(c.type().id());
  s.scale     = c.type().scale();
  s.size      = c.size();
  s.null_mask = const_cast<uint32_t*>(c.null_mask());
  if (c.type().id() == cudf::type_id::STRING) {
    s.data    = const_cast<char*>(c.head<char>());                       // chars live in the parent's data (RC:1919-1923)
    s.offsets = c.num_children() > 0 ? const_cast<int32_t*>(c.child(0).head<int32_t>()) : nullptr;
  } else {
    s.data = const_cast<uint8_t*>(c.head<uint8_t>());
  }
  return s;
}

inline int32_t size_of_type(int32_t t)
{
  srj_layout l{};
  int32_t st = 0, sz = 0;
  return srj_compute_layout(&t, 1, &l, &st, &sz) == SRJ_OK ? sz : 0;
}

// blocking device -> host copy on `stream` (cudaMemcpyAsync + synchronize in the real build)
bool copy_to_host(void* dst, const void* src, size_t bytes, rmm::cuda_stream_view stream);
bool copy_from_host(void* dst, const void* src, size_t bytes, rmm::cuda_stream_view stream);

template <typename T>
jlong release_as_jlong(std::unique_ptr<T>&& p) { return reinterpret_cast<jlong>(p.release()); }   // jni_utils.hpp:34-46

}  // namespace srjshim
