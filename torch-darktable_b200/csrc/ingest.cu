// Demosaic straight from the 12-bit packed frame: unpack + black level + white balance happen while the CFA patch is
// staged into shared memory (cfa_tile.cuh), so no float CFA plane ever exists in HBM.
// Replaces decode12_float -> apply_white_balance -> demosaic of pipeline/image_processor.py:190-247
// (5.5 + 16 + 16 B/px of traffic before the demosaic even starts) by 1.5 B in + 12 B out per pixel.
#include "launchers.cuh"

using namespace tdb;

extern "C" int tdb_demosaic_packed(const uint8_t *packed, float *rgb, int width, int height, int ids_format, uint32_t filters,
                                   int method, float black, const float *gains, float ppg_median_threshold, tdb_stream_t stream) {
  TDB_REQUIRE(packed && rgb, "demosaic_packed: null pointer");
  TDB_REQUIRE(width >= 16 && height >= 16 && !(width & 1) && !(height & 1),
              "demosaic_packed: width and height must be even and >= 16 (got %dx%d)", width, height);
  CfaSource src{};
  src.packed = packed;
  src.ids = ids_format;
  src.black = black;
  src.gains_dev = gains;
  src.apply = gains ? 2 : (black != 0.0f ? 1 : 0);
  cudaStream_t s = as_stream(stream);
  switch (method) {
    case TDB_DEMOSAIC_BILINEAR: return launch_bilinear(src, rgb, width, height, filters, s);
    case TDB_DEMOSAIC_PPG: return launch_ppg(src, rgb, width, height, filters, ppg_median_threshold, s);
    case TDB_DEMOSAIC_RCD: return launch_rcd(src, rgb, width, height, filters, s);
  }
  set_error("demosaic_packed: unknown method %d", method);
  return TDB_EINVAL;
}
