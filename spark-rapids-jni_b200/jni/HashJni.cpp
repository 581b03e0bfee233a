// HashJni.cpp -- JNI binding of com.nvidia.spark.rapids.jni.Hash over libsrj_b200.so.  Replaces
// src/main/cpp/src/hash/HashJni.cpp:26-79 of the reference: getMaxStackDepth, murmurHash32, xxhash64, hiveHash.
// Inputs: jlongArray of cudf::column_view*; output: a heap cudf::column* (INT32 / INT64, no null mask).
#include "srj_jni_common.hpp"

using namespace srjshim;

namespace {

enum class Kind { MURMUR, XXHASH64, HIVE };

jlong row_hash(JNIEnv* env, Kind kind, jlong seed, jlongArray column_handles)
{
  if (!column_handles) { throw_java(env, "java/lang/NullPointerException", "array of column handles is null"); return 0; }   // HashJni.cpp:36
  cudf::jni::auto_set_device(env);
  const int nc = env->GetArrayLength(column_handles);
  std::vector<srj_column> cols(nc);
  int64_t n = 0;
  {
    jlong* h = env->GetLongArrayElements(column_handles, nullptr);
    for (int c = 0; c < nc; ++c) {
      auto const* v = reinterpret_cast<cudf::column_view const*>(h[c]);
      if (!v) { env->ReleaseLongArrayElements(column_handles, h, JNI_ABORT); throw_java(env, "java/lang/NullPointerException", "column handle is null"); return 0; }
      cols[c] = to_srj(*v);
      n       = v->size();
    }
    env->ReleaseLongArrayElements(column_handles, h, JNI_ABORT);
  }
  auto stream      = cudf::get_default_stream();
  const bool wide  = kind == Kind::XXHASH64;
  rmm::device_buffer out(static_cast<size_t>(n) * (wide ? 8 : 4), stream);
  int st;
  switch (kind) {
    case Kind::MURMUR: st = srj_murmur_hash3_32(cols.data(), nc, n, static_cast<uint32_t>(seed), static_cast<int32_t*>(out.data()), stream.value()); break;
    case Kind::XXHASH64: st = srj_xxhash64(cols.data(), nc, n, seed, static_cast<int64_t*>(out.data()), stream.value()); break;
    default: st = srj_hive_hash(cols.data(), nc, n, static_cast<int32_t*>(out.data()), stream.value()); break;
  }
  if (throw_if_error(env, st)) return 0;
  return release_as_jlong(std::make_unique<cudf::column>(cudf::data_type{wide ? cudf::type_id::INT64 : cudf::type_id::INT32},
                                                         static_cast<cudf::size_type>(n), std::move(out), rmm::device_buffer{}, 0));
}

}  // namespace

extern "C" {

JNIEXPORT jint JNICALL Java_com_nvidia_spark_rapids_jni_Hash_getMaxStackDepth(JNIEnv*, jclass) { return srj_get_max_stack_depth(); }

JNIEXPORT jlong JNICALL Java_com_nvidia_spark_rapids_jni_Hash_murmurHash32(JNIEnv* env, jclass, jint seed, jlongArray column_handles)
{
  return row_hash(env, Kind::MURMUR, seed, column_handles);
}

JNIEXPORT jlong JNICALL Java_com_nvidia_spark_rapids_jni_Hash_xxhash64(JNIEnv* env, jclass, jlong seed, jlongArray column_handles)
{
  return row_hash(env, Kind::XXHASH64, seed, column_handles);
}

JNIEXPORT jlong JNICALL Java_com_nvidia_spark_rapids_jni_Hash_hiveHash(JNIEnv* env, jclass, jlongArray column_handles)
{
  return row_hash(env, Kind::HIVE, 0, column_handles);
}

}  // extern "C"
