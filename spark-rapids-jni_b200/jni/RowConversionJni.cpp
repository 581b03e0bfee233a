// RowConversionJni.cpp -- JNI binding of com.nvidia.spark.rapids.jni.RowConversion over libsrj_b200.so.
// Replaces src/main/cpp/src/RowConversionJni.cpp:23-124 of the reference: same four symbols, same handle ABI
// (jlong = cudf::table_view const* / cudf::column_view*; results are heap cudf::column* owned by Java), same
// exception classes.  Every device buffer is an rmm::device_buffer allocated here, exactly where the reference
// allocates (RC:1905-1923, 2220-2241, 2421-2428); the C ABI never allocates.
#include "srj_jni_common.hpp"

using namespace srjshim;

namespace {

// convertToRows / convertToRowsFixedWidthOptimized: Table -> one LIST<INT8> column per <= 2 GiB batch
jlongArray to_rows(JNIEnv* env, jlong input_table, bool fixed_only)
{
  if (input_table == 0) { throw_java(env, "java/lang/NullPointerException", "input table is null"); return nullptr; }   // JNI_NULL_CHECK
  cudf::jni::auto_set_device(env);
  auto const* tbl = reinterpret_cast<cudf::table_view const*>(input_table);
  auto stream     = cudf::get_default_stream();
  const int nc    = tbl->num_columns();
  const int64_t n = tbl->num_rows();
  std::vector<int32_t> types(nc), scales(nc);
  std::vector<srj_column> cols(nc);
  for (int c = 0; c < nc; ++c) {
    cols[c]   = to_srj(tbl->column(c));
    types[c]  = cols[c].type_id;
    scales[c] = cols[c].scale;
    if (fixed_only && types[c] == SRJ_STRING) { throw_java(env, "ai/rapids/cudf/CudfException", "Only fixed width types are currently supported"); return nullptr; }   // RC:2122-2124
  }
  int st;
  const srj_plan* plan = plan_for(types, scales, &st);
  if (throw_if_error(env, st)) return nullptr;
  rmm::device_buffer ws(static_cast<size_t>(srj_to_rows_workspace_bytes(plan, n)), stream);
  std::vector<srj_row_batch> batches(4096);
  int32_t nb = 0;
  if (throw_if_error(env, srj_to_rows_plan_batches(plan, cols.data(), n, ws.data(), batches.data(), 4096, &nb, stream.value()))) return nullptr;
  if (nb == 0) { nb = 1; batches[0] = srj_row_batch{0, 0, 0}; }          // empty table: one empty LIST column
  std::vector<std::unique_ptr<cudf::column>> offs(nb), data(nb);
  std::vector<int32_t*> optr(nb);
  std::vector<uint8_t*> dptr(nb);
  for (int b = 0; b < nb; ++b) {
    offs[b] = std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT32}, static_cast<cudf::size_type>(batches[b].row_count + 1),
                                             rmm::device_buffer(static_cast<size_t>(batches[b].row_count + 1) * 4, stream), rmm::device_buffer{}, 0);
    data[b] = std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT8}, static_cast<cudf::size_type>(batches[b].num_bytes),
                                             rmm::device_buffer(static_cast<size_t>(batches[b].num_bytes), stream), rmm::device_buffer{}, 0);
    optr[b] = offs[b]->mutable_view().head<int32_t>();
    dptr[b] = reinterpret_cast<uint8_t*>(data[b]->mutable_view().head<int8_t>());
  }
  if (n > 0 &&
      throw_if_error(env, srj_convert_to_rows(plan, cols.data(), n, ws.data(), batches.data(), nb, optr.data(), dptr.data(), stream.value())))
    return nullptr;
  std::vector<jlong> handles(nb);
  for (int b = 0; b < nb; ++b)
    handles[b] = release_as_jlong(cudf::make_lists_column(static_cast<cudf::size_type>(batches[b].row_count), std::move(offs[b]),
                                                          std::move(data[b]), 0, rmm::device_buffer{}));      // RC:1954-1979
  stream.synchronize();                                                   // ws is released on return
  jlongArray out = env->NewLongArray(nb);
  if (out) env->SetLongArrayRegion(out, 0, nb, handles.data());
  return out;
}

// convertFromRows / convertFromRowsFixedWidthOptimized: LIST<INT8> rows + schema -> Table
jlongArray from_rows(JNIEnv* env, jlong input_column, jintArray jtypes, jintArray jscales, bool fixed_only)
{
  if (input_column == 0 || !jtypes || !jscales) { throw_java(env, "java/lang/NullPointerException", "input is null"); return nullptr; }
  cudf::jni::auto_set_device(env);
  const int nc = env->GetArrayLength(jtypes);
  if (nc != env->GetArrayLength(jscales)) { throw_java(env, "java/lang/IllegalArgumentException", "types and scales must match size"); return nullptr; }   // RowConversionJni.cpp:80-83
  std::vector<int32_t> types(nc), scales(nc);
  {
    jint* t = env->GetIntArrayElements(jtypes, nullptr);
    jint* s = env->GetIntArrayElements(jscales, nullptr);
    for (int c = 0; c < nc; ++c) { types[c] = t[c]; scales[c] = s[c]; }
    env->ReleaseIntArrayElements(jtypes, t, JNI_ABORT);
    env->ReleaseIntArrayElements(jscales, s, JNI_ABORT);
  }
  auto const* lv = reinterpret_cast<cudf::column_view const*>(input_column);
  if (lv->type().id() != cudf::type_id::LIST || lv->num_children() < 2 ||
      (lv->child(1).type().id() != cudf::type_id::INT8 && lv->child(1).type().id() != cudf::type_id::UINT8)) {
    throw_java(env, "ai/rapids/cudf/CudfException", "Only a list of bytes is supported as input");                // RC:2157-2158
    return nullptr;
  }
  bool any_string = false;
  for (int c = 0; c < nc; ++c) any_string |= types[c] == SRJ_STRING;
  if (fixed_only && any_string) { throw_java(env, "ai/rapids/cudf/CudfException", "Only fixed width types are currently supported"); return nullptr; }   // RC:2509-2511
  auto stream              = cudf::get_default_stream();
  const int64_t n          = lv->size();
  const int32_t* d_offsets = lv->child(0).head<int32_t>();
  const uint8_t* d_rows    = lv->child(1).head<uint8_t>();
  const int64_t rows_bytes = lv->child(1).size();
  int st;
  const srj_plan* plan = plan_for(types, scales, &st);
  if (throw_if_error(env, st)) return nullptr;

  const size_t words = static_cast<size_t>((n + 31) / 32);
  std::vector<rmm::device_buffer> data(nc), masks(nc), offs(nc);
  std::vector<srj_column> cols(nc);
  for (int c = 0; c < nc; ++c) {
    masks[c] = rmm::device_buffer(words * 4, stream);                                                                // always allocated, RC:2220
    cols[c]  = srj_column{types[c], scales[c], n, nullptr, static_cast<uint32_t*>(masks[c].data()), nullptr};
    if (types[c] == SRJ_STRING) {
      offs[c]         = rmm::device_buffer(static_cast<size_t>(n + 1) * 4, stream);
      cols[c].offsets = static_cast<int32_t*>(offs[c].data());
    } else {
      data[c]      = rmm::device_buffer(static_cast<size_t>(n) * size_of_type(types[c]), stream);
      cols[c].data = data[c].data();
    }
  }
  rmm::device_buffer counters(static_cast<size_t>(2 * nc + 1) * 8, stream);   // null counts [nc] | char totals [nc + 1]
  auto* d_nulls  = static_cast<int64_t*>(counters.data());
  auto* d_totals = d_nulls + nc;
  rmm::device_buffer ws(static_cast<size_t>(srj_from_rows_workspace_bytes(plan, n)), stream);
  if (throw_if_error(env, srj_convert_from_rows_fixed(plan, d_rows, d_offsets, rows_bytes, n, cols.data(), d_nulls, d_totals, nullptr,
                                                      ws.data(), stream.value())))
    return nullptr;
  std::vector<int64_t> h(static_cast<size_t>(2 * nc + 1));
  if (!copy_to_host(h.data(), counters.data(), h.size() * 8, stream)) { throw_java(env, "ai/rapids/cudf/CudaException", "D2H of the counters failed"); return nullptr; }   // the sync of RC:2389
  if (any_string) {
    if (h[2 * nc] & 2) { throw_java(env, "ai/rapids/cudf/CudfColumnSizeOverflowException", "string column exceeds the int32 chars limit"); return nullptr; }
    for (int c = 0; c < nc; ++c)
      if (types[c] == SRJ_STRING) {
        data[c]      = rmm::device_buffer(static_cast<size_t>(h[nc + c]), stream);
        cols[c].data = data[c].data();
      }
    if (throw_if_error(env, srj_convert_from_rows_strings(plan, d_rows, d_offsets, rows_bytes, n, cols.data(), d_totals, ws.data(), stream.value())))
      return nullptr;
  }
  std::vector<jlong> handles(nc);
  for (int c = 0; c < nc; ++c) {
    const auto nulls = static_cast<cudf::size_type>(h[c]);
    if (types[c] == SRJ_STRING) {
      auto o     = std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT32}, static_cast<cudf::size_type>(n + 1), std::move(offs[c]), rmm::device_buffer{}, 0);
      handles[c] = release_as_jlong(cudf::make_strings_column(static_cast<cudf::size_type>(n), std::move(o), std::move(data[c]), nulls, std::move(masks[c])));   // RC:2421-2428
    } else {
      handles[c] = release_as_jlong(std::make_unique<cudf::column>(cudf::data_type{static_cast<cudf::type_id>(types[c]), scales[c]},
                                                                   static_cast<cudf::size_type>(n), std::move(data[c]), std::move(masks[c]), nulls));
    }
  }
  stream.synchronize();   // ws / counters are released on return
  jlongArray out = env->NewLongArray(nc);
  if (out) env->SetLongArrayRegion(out, 0, nc, handles.data());
  return out;
}

}  // namespace

extern "C" {

JNIEXPORT jlongArray JNICALL Java_com_nvidia_spark_rapids_jni_RowConversion_convertToRows(JNIEnv* env, jclass, jlong input_table)
{
  return to_rows(env, input_table, false);
}

JNIEXPORT jlongArray JNICALL Java_com_nvidia_spark_rapids_jni_RowConversion_convertToRowsFixedWidthOptimized(JNIEnv* env, jclass, jlong input_table)
{
  return to_rows(env, input_table, true);
}

JNIEXPORT jlongArray JNICALL Java_com_nvidia_spark_rapids_jni_RowConversion_convertFromRows(JNIEnv* env, jclass, jlong input_column,
                                                                                             jintArray types, jintArray scale)
{
  return from_rows(env, input_column, types, scale, false);
}

JNIEXPORT jlongArray JNICALL Java_com_nvidia_spark_rapids_jni_RowConversion_convertFromRowsFixedWidthOptimized(JNIEnv* env, jclass, jlong input_column,
                                                                                                                jintArray types, jintArray scale)
{
  return from_rows(env, input_column, types, scale, true);
}

}  // extern "C"
