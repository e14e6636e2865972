// Arrow C Device Data Interface hand-off (include/arrowgpu_cdata.h): export device-resident
// buffers as ArrowDeviceArray{ARROW_DEVICE_CUDA} with a cudaEvent_t* sync_event, import the same
// from any other producer.  No kernels here: the point is that record batches cross the
// boundary by pointer and stay in HBM (the reference's own device-array
// producers are CPU-only, arrow/cdata/exports.go:316,357).
#include "common.cuh"

#include "../../include/arrowgpu_cdata.h"

#include <stdlib.h>
#include <string.h>

namespace ag {
namespace {

struct ExportState {
  const void* buffers[2];
  cudaEvent_t event;
  bool have_event;
  void (*release_buffers)(void*);
  void* opaque;
};

void release_exported_array(struct ArrowArray* a) {
  if (!a || !a->release) return;
  ExportState* st = static_cast<ExportState*>(a->private_data);
  if (st) {
    if (st->have_event) cudaEventDestroy(st->event);
    if (st->release_buffers) st->release_buffers(st->opaque);
    free(st);
  }
  a->release = nullptr;  // the ABI's "released" marker
  a->private_data = nullptr;
}

void release_exported_schema(struct ArrowSchema* s) {
  if (!s || !s->release) return;
  s->release = nullptr;  // format / name point at static strings: nothing to free
}

}  // namespace
}  // namespace ag

using namespace ag;

extern "C" const char* ag_type_to_schema_format(int type) {
  switch (type) {
    case AG_TYPE_BOOL: return "b";
    case AG_TYPE_INT8: return "c";
    case AG_TYPE_UINT8: return "C";
    case AG_TYPE_INT16: return "s";
    case AG_TYPE_UINT16: return "S";
    case AG_TYPE_INT32: return "i";
    case AG_TYPE_UINT32: return "I";
    case AG_TYPE_INT64: return "l";
    case AG_TYPE_UINT64: return "L";
    case AG_TYPE_FLOAT32: return "f";
    case AG_TYPE_FLOAT64: return "g";
    default: return nullptr;
  }
}

extern "C" ag_status ag_schema_format_to_type(const char* format, int* type) {
  if (!format || !type) AG_FAIL(AG_ERR_INVALID, "cdata: NULL format / type");
  static const int ids[] = {AG_TYPE_BOOL, AG_TYPE_INT8, AG_TYPE_UINT8, AG_TYPE_INT16, AG_TYPE_UINT16, AG_TYPE_INT32,
                            AG_TYPE_UINT32, AG_TYPE_INT64, AG_TYPE_UINT64, AG_TYPE_FLOAT32, AG_TYPE_FLOAT64};
  for (int id : ids) {
    if (strcmp(format, ag_type_to_schema_format(id)) == 0) { *type = id; return AG_OK; }
  }
  AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "cdata: format '%s' is not a boolean / numeric primitive type", format);
}

extern "C" ag_status ag_device_array_describe(const struct ArrowDeviceArray* in, const struct ArrowSchema* schema, ag_array_view* out) {
  if (!in || !out) AG_FAIL(AG_ERR_INVALID, "cdata: NULL array / view");
  const struct ArrowArray& a = in->array;
  if (!a.release) AG_FAIL(AG_ERR_INVALID, "cdata: array was already released");
  if (a.n_buffers != 2) AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "cdata: expected 2 buffers (validity, values), got %lld", (long long)a.n_buffers);
  if (a.n_children != 0 || a.dictionary) AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "cdata: nested / dictionary arrays are outside this path");
  if (a.length < 0 || a.offset < 0) AG_FAIL(AG_ERR_INVALID, "cdata: negative length or offset");
  if (!a.buffers) AG_FAIL(AG_ERR_INVALID, "cdata: NULL buffers");
  out->type = 0;
  if (schema) {
    if (!schema->release) AG_FAIL(AG_ERR_INVALID, "cdata: schema was already released");
    AG_TRY(ag_schema_format_to_type(schema->format, &out->type));
  }
  out->length = a.length;
  out->null_count = a.null_count;
  out->offset = a.offset;
  out->validity = static_cast<const uint8_t*>(a.buffers[0]);
  out->values = a.buffers[1];
  out->device_type = in->device_type;
  out->device_id = in->device_id;
  if (a.length > 0 && !out->values) AG_FAIL(AG_ERR_INVALID, "cdata: NULL values buffer");
  return AG_OK;
}

extern "C" ag_status ag_export_device_array(int type, int64_t length, int64_t null_count, int64_t offset,
                                            const void* d_validity, const void* d_values,
                                            void (*release_buffers)(void*), void* opaque, ag_stream_t produced_on,
                                            struct ArrowDeviceArray* out, struct ArrowSchema* out_schema) {
  AG_TRY(ensure_init());
  if (!out) AG_FAIL(AG_ERR_INVALID, "cdata: NULL output");
  const char* fmt = ag_type_to_schema_format(type);
  if (!fmt) AG_FAIL(AG_ERR_TYPE, "cdata: unsupported type id %d", type);
  if (length < 0 || offset < 0) AG_FAIL(AG_ERR_INVALID, "cdata: negative length or offset");
  if (length > 0 && !d_values) AG_FAIL(AG_ERR_INVALID, "cdata: NULL values buffer");
  ExportState* st = static_cast<ExportState*>(calloc(1, sizeof(ExportState)));
  if (!st) AG_FAIL(AG_ERR_OOM, "cdata: out of host memory");
  st->buffers[0] = d_validity;
  st->buffers[1] = d_values;
  st->release_buffers = release_buffers;
  st->opaque = opaque;
  cudaError_t e = cudaEventCreateWithFlags(&st->event, cudaEventDisableTiming);
  if (e == cudaSuccess) {
    st->have_event = true;
    e = cudaEventRecord(st->event, resolve_stream(produced_on));
  }
  if (e != cudaSuccess) {
    if (st->have_event) cudaEventDestroy(st->event);
    free(st);
    AG_FAIL(AG_ERR_CUDA, "cdata: %s", cudaGetErrorString(e));
  }
  int dev = 0;
  cudaGetDevice(&dev);
  memset(out, 0, sizeof(*out));
  out->array.length = length;
  out->array.null_count = null_count;
  out->array.offset = offset;
  out->array.n_buffers = 2;
  out->array.n_children = 0;
  out->array.buffers = st->buffers;
  out->array.children = nullptr;
  out->array.dictionary = nullptr;
  out->array.release = release_exported_array;
  out->array.private_data = st;
  out->device_id = dev;
  out->device_type = ARROW_DEVICE_CUDA;
  out->sync_event = &st->event;
  if (out_schema) {
    memset(out_schema, 0, sizeof(*out_schema));
    out_schema->format = fmt;
    out_schema->name = "";
    out_schema->metadata = nullptr;
    out_schema->flags = ARROW_FLAG_NULLABLE;
    out_schema->release = release_exported_schema;
  }
  return AG_OK;
}

extern "C" ag_status ag_import_device_array(struct ArrowDeviceArray* in, const struct ArrowSchema* schema, ag_stream_t consume_on, ag_array_view* out) {
  AG_TRY(ensure_init());
  AG_TRY(ag_device_array_describe(in, schema, out));
  switch (in->device_type) {
    case ARROW_DEVICE_CUDA: {
      int dev = 0;
      cudaGetDevice(&dev);
      if (in->device_id != dev)
        AG_FAIL(AG_ERR_INVALID, "cdata: array lives on CUDA device %lld, the calling thread works on device %d (ag_set_device first)", (long long)in->device_id, dev);
      break;
    }
    case ARROW_DEVICE_CUDA_HOST:
    case ARROW_DEVICE_CUDA_MANAGED:
      break;  // pinned / managed memory: addressable from the kernels through unified addressing
    case ARROW_DEVICE_CPU:
      AG_FAIL(AG_ERR_INVALID, "cdata: ARROW_DEVICE_CPU array — pageable host memory is not device-addressable; upload it (ag_upload) or allocate it with ag_host_alloc and export it as ARROW_DEVICE_CUDA_HOST");
    default:
      AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "cdata: device type %d is not a CUDA device", (int)in->device_type);
  }
  if (in->sync_event) {
    cudaEvent_t ev = *static_cast<cudaEvent_t*>(in->sync_event);
    AG_CUDA_TRY(cudaStreamWaitEvent(resolve_stream(consume_on), ev, 0));
  }
  return AG_OK;
}
