// placeholder until the tcgen05 kNN kernel lands
#include "ssw_db.h"
extern "C" {
int ssw_knn_build(int, const void*, int, int64_t, int, int, int64_t, int64_t, int32_t*, float*) {
  ssw::set_error("ssw_knn_build: not built yet");
  return SSW_ERR_INVALID;
}
int ssw_knn_build_device(int, const void*, int64_t, int, int, int64_t, int64_t, int32_t*, float*, void*) {
  ssw::set_error("ssw_knn_build_device: not built yet");
  return SSW_ERR_INVALID;
}
}
