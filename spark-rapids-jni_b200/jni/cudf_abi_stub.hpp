// cudf_abi_stub.hpp -- stand-ins for the libcudf / rmm declarations the shim touches, so that it can be
// syntax-checked without libcudf (this image has neither libcudf nor rmm headers).  Field meanings follow
// thirdparty/cudf/cpp/include/cudf/column/column_view.hpp:237-244 and types.hpp:191-224; a real build includes
// <cudf/column/column_view.hpp>, <cudf/table/table_view.hpp>, <cudf/column/column_factories.hpp>,
// <cudf/lists/lists_column_view.hpp>, <rmm/device_buffer.hpp>, <rmm/cuda_stream_view.hpp> instead.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace rmm {
struct cuda_stream_view { void* value() const; void synchronize() const; };
struct device_buffer {
  device_buffer();
  device_buffer(std::size_t bytes, cuda_stream_view stream);
  void* data();
  std::size_t size() const;
};
}  // namespace rmm

namespace cudf {
using size_type     = int32_t;
using bitmask_type  = uint32_t;
enum class type_id : int32_t { EMPTY = 0, INT8 = 1, UINT8 = 5, INT32 = 3, INT64 = 4, STRING = 23, LIST = 24 };
struct data_type {
  data_type(type_id id, int32_t scale = 0);
  type_id id() const;
  int32_t scale() const;
};
struct column_view {
  column_view(data_type type, size_type size, void const* data, bitmask_type const* null_mask, size_type null_count, size_type offset = 0,
              std::vector<column_view> const& children = {});
  data_type type() const;
  size_type size() const;
  size_type offset() const;
  size_type num_children() const;
  column_view child(size_type i) const;
  bitmask_type const* null_mask() const;
  template <typename T> T const* head() const;     // base pointer, offset not applied
};
struct mutable_column_view : column_view {
  template <typename T> T* head() const;
  bitmask_type* null_mask() const;
};
struct table_view {
  size_type num_columns() const;
  size_type num_rows() const;
  column_view column(size_type i) const;
};
struct column {
  column(data_type type, size_type size, rmm::device_buffer&& data, rmm::device_buffer&& null_mask, size_type null_count);
  mutable_column_view mutable_view();
  void set_null_count(size_type n);
};
std::unique_ptr<column> make_lists_column(size_type num_rows, std::unique_ptr<column> offsets, std::unique_ptr<column> child,
                                          size_type null_count, rmm::device_buffer&& null_mask);
std::unique_ptr<column> make_strings_column(size_type num_rows, std::unique_ptr<column> offsets, rmm::device_buffer&& chars,
                                            size_type null_count, rmm::device_buffer&& null_mask);
rmm::cuda_stream_view get_default_stream();
namespace jni {
void auto_set_device(JNIEnv* env);
}  // namespace jni
}  // namespace cudf
