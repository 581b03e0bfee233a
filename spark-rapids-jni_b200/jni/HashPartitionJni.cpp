// HashPartitionJni.cpp -- JNI binding of a com.nvidia.spark.rapids.jni.HashPartition class (new: the reference leaves
// this step to the plugin's GpuHashPartitioning + ai.rapids.cudf.Table.partition) over libsrj_b200.so:
//   static native long[] hashPartition(long tableView, int[] keyColumns, int numPartitions, int seed)
// returns { INT32 column of the P partition start rows, then one column handle per table column } -- the pieces of an
// ai.rapids.cudf.PartitionedTable; rows of a partition keep their input order.
#include "srj_jni_common.hpp"

using namespace srjshim;

extern "C" {

JNIEXPORT jlongArray JNICALL Java_com_nvidia_spark_rapids_jni_HashPartition_hashPartition(JNIEnv* env, jclass, jlong j_table_view, jintArray j_keys,
                                                                                         jint num_partitions, jint seed)
{
  if (!j_table_view || !j_keys) { throw_java(env, "java/lang/NullPointerException", "table / key columns is null"); return nullptr; }
  cudf::jni::auto_set_device(env);
  auto const* tbl = reinterpret_cast<cudf::table_view const*>(j_table_view);
  auto stream     = cudf::get_default_stream();
  const int nc    = tbl->num_columns();
  const int64_t n = tbl->num_rows();
  std::vector<srj_column> cols(nc), keys;
  for (int c = 0; c < nc; ++c) cols[c] = to_srj(tbl->column(c));
  {
    const int nk = env->GetArrayLength(j_keys);
    jint* h      = env->GetIntArrayElements(j_keys, nullptr);
    bool ok      = nk > 0;
    for (int k = 0; k < nk && ok; ++k) {
      ok = h[k] >= 0 && h[k] < nc;
      if (ok) keys.push_back(cols[h[k]]);
    }
    env->ReleaseIntArrayElements(j_keys, h, JNI_ABORT);
    if (!ok) { throw_java(env, "java/lang/IllegalArgumentException", "key column index out of range"); return nullptr; }
  }
  const int P = num_partitions;
  rmm::device_buffer ws(static_cast<size_t>(srj_partition_workspace_bytes(n, P)), stream);
  rmm::device_buffer ids(static_cast<size_t>(n) * 4, stream), smap(static_cast<size_t>(n) * 4, stream), gmap(static_cast<size_t>(n) * 4, stream);
  rmm::device_buffer offsets(static_cast<size_t>(P + 1) * 4, stream);
  if (throw_if_error(env, srj_hash_partition(keys.data(), static_cast<int32_t>(keys.size()), n, static_cast<uint32_t>(seed), P, static_cast<int32_t*>(ids.data()),
                                             static_cast<int32_t*>(offsets.data()), static_cast<int32_t*>(smap.data()), static_cast<int32_t*>(gmap.data()), ws.data(),
                                             stream.value())))
    return nullptr;
  // outputs: same types and sizes as the inputs
  const size_t mask_bytes = static_cast<size_t>((n + 31) / 32) * 4;
  std::vector<std::unique_ptr<cudf::column>> out_offs(nc);
  std::vector<rmm::device_buffer> masks(nc), bufs(nc);
  std::vector<srj_column> out(nc);
  for (int c = 0; c < nc; ++c) {
    out[c] = cols[c];
    if (cols[c].null_mask) {
      masks[c]         = rmm::device_buffer(mask_bytes, stream);
      out[c].null_mask = static_cast<uint32_t*>(masks[c].data());
    }
    if (cols[c].type_id == SRJ_STRING) {
      out_offs[c]    = std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT32}, static_cast<cudf::size_type>(n + 1),
                                                      rmm::device_buffer(static_cast<size_t>(n + 1) * 4, stream), rmm::device_buffer{}, 0);
      out[c].offsets = out_offs[c]->mutable_view().head<int32_t>();
      out[c].data    = nullptr;
    } else {
      bufs[c]     = rmm::device_buffer(static_cast<size_t>(n) * size_of_type(cols[c].type_id), stream);
      out[c].data = bufs[c].data();
    }
  }
  rmm::device_buffer d_nulls(static_cast<size_t>(nc) * 8, stream);
  if (throw_if_error(env, srj_partition_columns(cols.data(), out.data(), nc, n, P, static_cast<int32_t*>(smap.data()), static_cast<int32_t*>(gmap.data()),
                                                static_cast<int64_t*>(d_nulls.data()), ws.data(), stream.value())))
    return nullptr;
  std::vector<int64_t> nulls(nc, 0);
  if (nc && !copy_to_host(nulls.data(), d_nulls.data(), static_cast<size_t>(nc) * 8, stream)) { throw_java(env, "ai/rapids/cudf/CudaException", "copy of the null counts failed"); return nullptr; }
  // STRING columns: chars sized by the last offset, then the second call
  std::vector<rmm::device_buffer> chars(nc);
  bool any_string = false;
  for (int c = 0; c < nc; ++c) {
    if (cols[c].type_id != SRJ_STRING) continue;
    int32_t total = 0;
    if (n > 0 && !copy_to_host(&total, out[c].offsets + n, 4, stream)) { throw_java(env, "ai/rapids/cudf/CudaException", "copy of a chars total failed"); return nullptr; }
    chars[c]    = rmm::device_buffer(static_cast<size_t>(total), stream);
    out[c].data = chars[c].data();
    any_string  = true;
  }
  if (any_string && throw_if_error(env, srj_partition_strings(cols.data(), out.data(), nc, n, static_cast<int32_t*>(gmap.data()), stream.value()))) return nullptr;
  stream.synchronize();
  std::vector<jlong> handles(nc + 1);
  handles[0] = release_as_jlong(std::make_unique<cudf::column>(cudf::data_type{cudf::type_id::INT32}, static_cast<cudf::size_type>(P), std::move(offsets),
                                                               rmm::device_buffer{}, 0));   // P start rows (+ the row count behind them)
  for (int c = 0; c < nc; ++c) {
    const auto nn = static_cast<cudf::size_type>(nulls[c]);
    if (cols[c].type_id == SRJ_STRING) {
      handles[c + 1] = release_as_jlong(cudf::make_strings_column(static_cast<cudf::size_type>(n), std::move(out_offs[c]), std::move(chars[c]), nn, std::move(masks[c])));
    } else {
      handles[c + 1] = release_as_jlong(std::make_unique<cudf::column>(cudf::data_type{static_cast<cudf::type_id>(cols[c].type_id), cols[c].scale},
                                                                       static_cast<cudf::size_type>(n), std::move(bufs[c]), std::move(masks[c]), nn));
    }
  }
  jlongArray jout = env->NewLongArray(nc + 1);
  if (jout) env->SetLongArrayRegion(jout, 0, nc + 1, handles.data());
  return jout;
}

}  // extern "C"
