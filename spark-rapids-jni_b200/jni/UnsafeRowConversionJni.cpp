// UnsafeRowConversionJni.cpp -- JNI binding of a com.nvidia.spark.rapids.jni.UnsafeRowConversion class (new; shaped after
// RowConversion.java:120-174 / RowConversionJni.cpp:23-124) over libsrj_b200.so: columns <-> Apache Spark UnsafeRow bytes
// carried, like the JCUDF rows, as one LIST<INT8> column.
//   static native long convertToRows(long tableView);                                     -> LIST<INT8> column handle
//   static native long[] convertFromRows(long listColumnView, int[] types, int[] scales); -> column handles
#include "srj_jni_common.hpp"

using namespace srjshim;

extern "C" {

JNIEXPORT jlong JNICALL Java_com_nvidia_spark_rapids_jni_UnsafeRowConversion_convertToRows(JNIEnv* env, jclass, jlong j_table_view)
{
  if (!j_table_view) { throw_java(env, "java/lang/NullPointerException", "input table is null"); return 0; }
  cudf::jni::auto_set_device(env);
  auto const* tbl = reinterpret_cast<cudf::table_view const*>(j_table_view);
  auto stream     = cudf::get_default_stream();
  const int nc    = tbl->num_columns();
  const int64_t n = tbl->num_rows();
  std::vector<srj_column> cols(nc);
  for (int c = 0; c < nc; ++c) cols[c] = to_srj(tbl->column(c));
  rmm::device_buffer ws(static_cast<size_t>(srj_unsafe_row_workspace_bytes(nc, n)), stream);
  auto offs = std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT32}, static_cast<cudf::size_type>(n + 1),
                                             rmm::device_buffer(static_cast<size_t>(n + 1) * 4, stream), rmm::device_buffer{}, 0);
  int32_t* d_offs = offs->mutable_view().head<int32_t>();
  int64_t total   = 0;
  if (throw_if_error(env, srj_unsafe_row_sizes(cols.data(), nc, n, d_offs, &total, ws.data(), stream.value()))) return 0;   // SRJ_EOVERFLOW -> CudfColumnSizeOverflowException
  auto data = std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT8}, static_cast<cudf::size_type>(total),
                                             rmm::device_buffer(static_cast<size_t>(total), stream), rmm::device_buffer{}, 0);
  if (throw_if_error(env, srj_convert_to_unsafe_rows(cols.data(), nc, n, d_offs, reinterpret_cast<uint8_t*>(data->mutable_view().head<int8_t>()), ws.data(),
                                                     stream.value())))
    return 0;
  stream.synchronize();
  return release_as_jlong(cudf::make_lists_column(static_cast<cudf::size_type>(n), std::move(offs), std::move(data), 0, rmm::device_buffer{}));
}

JNIEXPORT jlongArray JNICALL Java_com_nvidia_spark_rapids_jni_UnsafeRowConversion_convertFromRows(JNIEnv* env, jclass, jlong j_list_view, jintArray j_types,
                                                                                                 jintArray j_scales)
{
  if (!j_list_view || !j_types || !j_scales) { throw_java(env, "java/lang/NullPointerException", "null argument"); return nullptr; }
  cudf::jni::auto_set_device(env);
  auto const* list = reinterpret_cast<cudf::column_view const*>(j_list_view);
  auto stream      = cudf::get_default_stream();
  const int nc     = env->GetArrayLength(j_types);
  if (env->GetArrayLength(j_scales) != nc) { throw_java(env, "java/lang/IllegalArgumentException", "types and scales must match size"); return nullptr; }   // RowConversionJni.cpp:80-83
  std::vector<int32_t> types(nc), scales(nc);
  {
    jint* t = env->GetIntArrayElements(j_types, nullptr);
    jint* s = env->GetIntArrayElements(j_scales, nullptr);
    for (int c = 0; c < nc; ++c) { types[c] = t[c]; scales[c] = s[c]; }
    env->ReleaseIntArrayElements(j_types, t, JNI_ABORT);
    env->ReleaseIntArrayElements(j_scales, s, JNI_ABORT);
  }
  const int64_t n        = list->size();
  const int32_t* d_offs  = list->child(0).head<int32_t>();
  const uint8_t* d_rows  = list->child(1).head<uint8_t>();
  const size_t mask_bytes = static_cast<size_t>((n + 31) / 32) * 4;
  rmm::device_buffer ws(static_cast<size_t>(srj_unsafe_row_workspace_bytes(nc, n)), stream);
  std::vector<std::unique_ptr<cudf::column>> offs(nc);
  std::vector<rmm::device_buffer> masks(nc), chars(nc), bufs(nc);
  std::vector<srj_column> out(nc);
  for (int c = 0; c < nc; ++c) {
    out[c]           = srj_column{};
    out[c].type_id   = types[c];
    out[c].scale     = scales[c];
    out[c].size      = n;
    masks[c]         = rmm::device_buffer(mask_bytes, stream);
    out[c].null_mask = static_cast<uint32_t*>(masks[c].data());
    if (types[c] == SRJ_STRING) {
      offs[c]        = std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT32}, static_cast<cudf::size_type>(n + 1),
                                                      rmm::device_buffer(static_cast<size_t>(n + 1) * 4, stream), rmm::device_buffer{}, 0);
      out[c].offsets = offs[c]->mutable_view().head<int32_t>();
    } else {
      bufs[c]     = rmm::device_buffer(static_cast<size_t>(n) * size_of_type(types[c]), stream);
      out[c].data = bufs[c].data();
    }
  }
  rmm::device_buffer d_nulls(static_cast<size_t>(nc) * 8, stream);
  if (throw_if_error(env, srj_convert_from_unsafe_rows(d_rows, d_offs, n, out.data(), nc, static_cast<int64_t*>(d_nulls.data()), ws.data(), stream.value()))) return nullptr;
  std::vector<int64_t> nulls(nc, 0);
  if (nc && !copy_to_host(nulls.data(), d_nulls.data(), static_cast<size_t>(nc) * 8, stream)) { throw_java(env, "ai/rapids/cudf/CudaException", "copy of the null counts failed"); return nullptr; }
  bool any_string = false;
  for (int c = 0; c < nc; ++c) {
    if (types[c] != SRJ_STRING) continue;
    int32_t total = 0;
    if (n > 0 && !copy_to_host(&total, out[c].offsets + n, 4, stream)) { throw_java(env, "ai/rapids/cudf/CudaException", "copy of a chars total failed"); return nullptr; }
    chars[c]    = rmm::device_buffer(static_cast<size_t>(total), stream);
    out[c].data = chars[c].data();
    any_string  = true;
  }
  if (any_string && throw_if_error(env, srj_convert_from_unsafe_rows_strings(d_rows, d_offs, n, out.data(), nc, stream.value()))) return nullptr;
  stream.synchronize();
  std::vector<jlong> handles(nc);
  for (int c = 0; c < nc; ++c) {
    const auto nn = static_cast<cudf::size_type>(nulls[c]);
    if (types[c] == SRJ_STRING) handles[c] = release_as_jlong(cudf::make_strings_column(static_cast<cudf::size_type>(n), std::move(offs[c]), std::move(chars[c]), nn, std::move(masks[c])));
    else
      handles[c] = release_as_jlong(std::make_unique<cudf::column>(cudf::data_type{static_cast<cudf::type_id>(types[c]), scales[c]}, static_cast<cudf::size_type>(n),
                                                                   std::move(bufs[c]), std::move(masks[c]), nn));
  }
  jlongArray jout = env->NewLongArray(nc);
  if (jout) env->SetLongArrayRegion(jout, 0, nc, handles.data());
  return jout;
}

}  // extern "C"
