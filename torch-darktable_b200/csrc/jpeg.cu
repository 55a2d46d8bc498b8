// JPEG output of the sRGB result (the step right after the hot path).
//
// Reference: csrc/jpeg_encoder.cu:104-180 -- one nvJPEG handle + encoder state per `Jpeg` object, a fresh parameter object per
// call (quality, optimised Huffman tables, chroma subsampling, baseline / progressive), nvjpegEncodeImage on torch's current
// stream and two nvjpegEncodeRetrieveBitstream calls (length, then bytes) into a fresh CPU tensor.
//
// Here nvJPEG stays what it is, a vendor library (the entropy coder is not a kernel of this repository), but
//   * it is bound lazily with dlopen, so libtdb200.so has no link-time dependency on it and every other entry point works
//     on a machine without libnvjpeg,
//   * the encoder is stream-ordered on the caller's stream: the uint8 RGBI image produced by tdb_tonemap /
//     tdb_bilateral_slice_tonemap (already rotated / flipped by those kernels) is consumed in place, with an explicit row
//     pitch, so no `.contiguous()` pass sits between the tone map and the encoder,
//   * the parameter object is cached per handle and only touched when a setting changes.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "tdb_common.cuh"

namespace tdb {
namespace {

struct Api {
  void *so = nullptr;
  decltype(&nvjpegCreateSimple) create = nullptr;
  decltype(&nvjpegDestroy) destroy = nullptr;
  decltype(&nvjpegEncoderStateCreate) state_create = nullptr;
  decltype(&nvjpegEncoderStateDestroy) state_destroy = nullptr;
  decltype(&nvjpegEncoderParamsCreate) params_create = nullptr;
  decltype(&nvjpegEncoderParamsDestroy) params_destroy = nullptr;
  decltype(&nvjpegEncoderParamsSetQuality) set_quality = nullptr;
  decltype(&nvjpegEncoderParamsSetEncoding) set_encoding = nullptr;
  decltype(&nvjpegEncoderParamsSetOptimizedHuffman) set_huffman = nullptr;
  decltype(&nvjpegEncoderParamsSetSamplingFactors) set_sampling = nullptr;
  decltype(&nvjpegEncodeImage) encode = nullptr;
  decltype(&nvjpegEncodeRetrieveBitstream) retrieve = nullptr;
  bool ok = false;
};

Api &api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      a.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.so) break;
    }
    if (!a.so) return;
    bool all = true;
    auto bind = [&](auto &fn, const char *sym) {
      fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(a.so, sym));
      all = all && fn != nullptr;
    };
    bind(a.create, "nvjpegCreateSimple");
    bind(a.destroy, "nvjpegDestroy");
    bind(a.state_create, "nvjpegEncoderStateCreate");
    bind(a.state_destroy, "nvjpegEncoderStateDestroy");
    bind(a.params_create, "nvjpegEncoderParamsCreate");
    bind(a.params_destroy, "nvjpegEncoderParamsDestroy");
    bind(a.set_quality, "nvjpegEncoderParamsSetQuality");
    bind(a.set_encoding, "nvjpegEncoderParamsSetEncoding");
    bind(a.set_huffman, "nvjpegEncoderParamsSetOptimizedHuffman");
    bind(a.set_sampling, "nvjpegEncoderParamsSetSamplingFactors");
    bind(a.encode, "nvjpegEncodeImage");
    bind(a.retrieve, "nvjpegEncodeRetrieveBitstream");
    a.ok = all;
  });
  return a;
}

// the texts of the reference's JpegException (jpeg_encoder.cu:52-66)
const char *status_text(int code) {
  switch (code) {
    case NVJPEG_STATUS_SUCCESS: return "success";
    case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
    case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
    case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
    case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "not supported";
    case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocation failed";
    case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
    case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
    case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
    default: return "unknown";
  }
}

struct Coder {
  nvjpegHandle_t handle = nullptr;
  nvjpegEncoderState_t state = nullptr;
  nvjpegEncoderParams_t params = nullptr;
  int quality = -1, subsampling = -1, progressive = -1;
  bool pending = false;  // an encoded image waits in `state`
};

int fail(const char *what, int code) {
  set_error("%s, nvjpeg error %d: %s", what, code, status_text(code));
  return TDB_EJPEG;
}

}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

int tdb_jpeg_available(void) { return api().ok ? 1 : 0; }

int tdb_jpeg_create(void **coder) {
  TDB_REQUIRE(coder, "jpeg_create: null pointer");
  *coder = nullptr;
  Api &n = api();
  if (!n.ok) {
    set_error("jpeg_create: libnvjpeg.so.12 could not be loaded (%s)", n.so ? "missing symbols" : "dlopen failed");
    return TDB_EUNSUPPORTED;
  }
  Coder *c = new Coder();
  int st = n.create(&c->handle);
  if (st == NVJPEG_STATUS_SUCCESS) st = n.state_create(c->handle, &c->state, nullptr);
  if (st == NVJPEG_STATUS_SUCCESS) st = n.params_create(c->handle, &c->params, nullptr);
  if (st != NVJPEG_STATUS_SUCCESS) {
    if (c->params) n.params_destroy(c->params);
    if (c->state) n.state_destroy(c->state);
    if (c->handle) n.destroy(c->handle);
    delete c;
    return fail("nvjpegCreateSimple", st);
  }
  *coder = c;
  return TDB_OK;
}

int tdb_jpeg_destroy(void *coder) {
  if (!coder) return TDB_OK;
  Coder *c = static_cast<Coder *>(coder);
  Api &n = api();
  n.params_destroy(c->params);
  n.state_destroy(c->state);
  n.destroy(c->handle);
  delete c;
  return TDB_OK;
}

int tdb_jpeg_encode(void *coder, const uint8_t *image, int width, int height, int64_t row_pitch, int64_t plane_stride, int input_format,
                    int quality, int subsampling, int progressive, size_t *length, tdb_stream_t stream) {
  TDB_REQUIRE(coder && image && length, "jpeg_encode: null pointer");
  TDB_REQUIRE(width > 0 && height > 0, "jpeg_encode: image dimensions must be positive");
  TDB_REQUIRE(quality >= 1 && quality <= 100, "jpeg_encode: quality must be in [1, 100], got %d", quality);
  Coder *c = static_cast<Coder *>(coder);
  Api &n = api();
  cudaStream_t s = as_stream(stream);

  nvjpegInputFormat_t fmt;
  bool interleaved = false;
  switch (input_format) {
    case TDB_JPEG_BGR: fmt = NVJPEG_INPUT_BGR; break;
    case TDB_JPEG_RGB: fmt = NVJPEG_INPUT_RGB; break;
    case TDB_JPEG_BGRI: fmt = NVJPEG_INPUT_BGRI, interleaved = true; break;
    case TDB_JPEG_RGBI: fmt = NVJPEG_INPUT_RGBI, interleaved = true; break;
    default: set_error("Invalid input format"); return TDB_EINVAL;
  }
  nvjpegChromaSubsampling_t css;
  switch (subsampling) {
    case TDB_JPEG_CSS_444: css = NVJPEG_CSS_444; break;
    case TDB_JPEG_CSS_422: css = NVJPEG_CSS_422; break;
    case TDB_JPEG_CSS_GRAY: css = NVJPEG_CSS_GRAY; break;
    default: set_error("Invalid subsampling"); return TDB_EINVAL;
  }
  TDB_REQUIRE(row_pitch >= (int64_t)width * (interleaved ? 3 : 1), "jpeg_encode: row pitch %lld shorter than a row", (long long)row_pitch);
  TDB_REQUIRE(interleaved || plane_stride >= row_pitch * height, "jpeg_encode: plane stride shorter than a plane");

  int st;
  if (c->quality != quality) {
    if ((st = n.set_quality(c->params, quality, s)) != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncoderParamsSetQuality", st);
    if (c->quality < 0 && (st = n.set_huffman(c->params, 1, s)) != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncoderParamsSetOptimizedHuffman", st);
    c->quality = quality;
  }
  if (c->subsampling != subsampling) {
    if ((st = n.set_sampling(c->params, css, s)) != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncoderParamsSetSamplingFactors", st);
    c->subsampling = subsampling;
  }
  if (c->progressive != (progressive ? 1 : 0)) {
    st = n.set_encoding(c->params, progressive ? NVJPEG_ENCODING_PROGRESSIVE_DCT_HUFFMAN : NVJPEG_ENCODING_BASELINE_DCT, s);
    if (st != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncoderParamsSetEncoding", st);
    c->progressive = progressive ? 1 : 0;
  }

  nvjpegImage_t img{};
  if (interleaved) {
    img.channel[0] = const_cast<unsigned char *>(image);
    img.pitch[0] = (size_t)row_pitch;
  } else {
    for (int i = 0; i < 3; i++) {
      img.channel[i] = const_cast<unsigned char *>(image) + plane_stride * i;
      img.pitch[i] = (size_t)row_pitch;
    }
  }
  c->pending = false;
  if ((st = n.encode(c->handle, c->state, c->params, &img, fmt, width, height, s)) != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeImage", st);
  size_t len = 0;
  if ((st = n.retrieve(c->handle, c->state, nullptr, &len, s)) != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeRetrieveBitstream", st);
  c->pending = true;
  *length = len;
  return TDB_OK;
}

int tdb_jpeg_retrieve(void *coder, uint8_t *host_out, size_t capacity, size_t *length, tdb_stream_t stream) {
  TDB_REQUIRE(coder && host_out && length, "jpeg_retrieve: null pointer");
  Coder *c = static_cast<Coder *>(coder);
  TDB_REQUIRE(c->pending, "jpeg_retrieve: no encoded image is waiting (call tdb_jpeg_encode first)");
  Api &n = api();
  size_t len = capacity;
  int st = n.retrieve(c->handle, c->state, nullptr, &len, as_stream(stream));
  if (st != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeRetrieveBitstream", st);
  TDB_REQUIRE(len <= capacity, "jpeg_retrieve: buffer of %zu bytes is too small for a %zu byte stream", capacity, len);
  if ((st = n.retrieve(c->handle, c->state, host_out, &len, as_stream(stream))) != NVJPEG_STATUS_SUCCESS)
    return fail("nvjpegEncodeRetrieveBitstream", st);
  // the bytes must be in host_out when this returns, whatever nvJPEG does internally on a non-default stream
  const cudaError_t err = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) {
    set_error("jpeg_retrieve: %s", cudaGetErrorString(err));
    return TDB_ECUDA;
  }
  *length = len;
  c->pending = false;
  return TDB_OK;
}

}  // extern "C"
