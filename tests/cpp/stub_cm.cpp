// stub_cm.cpp -- host-only stand-in for the few C-ABI entry points FusedFrame uses, for the ThreadSanitizer test of the
// shim's hand-off logic (tests/cpp/test_shim_threads.cpp). It computes nothing: it records which sensors were submitted
// and merged, under its own mutex like the real library (cm_api.cu takes the handle mutex in every entry point). CUDA
// cannot run under TSAN, and the logic under test -- the flag gate between six callback threads and the main loop -- is
// host code in include/cloud_merger_shim.hpp.
#include <cstring>
#include <mutex>

#include "cloud_merger_gpu.h"

struct cm_handle_s {
  std::mutex mu;
  uint64_t submitted = 0;
  int policy = CM_SUBMIT_LATEST_WINS;
  uint64_t last_used = 0;
  int64_t ticket = 0;
  long long submits = 0, dropped = 0, merges = 0;
};

extern "C" {
int cm_create(const cm_config_t*, cm_handle_t* out) { *out = new cm_handle_s(); return CM_OK; }
int cm_destroy(cm_handle_t h) { delete h; return CM_OK; }
const char* cm_strerror(int) { return "stub"; }
const char* cm_last_error(cm_handle_t) { return ""; }
int cm_set_crop(cm_handle_t, int, const cm_pass_t*) { return CM_OK; }
int cm_set_voxel(cm_handle_t, const float*, int, int) { return CM_OK; }
int cm_set_overflow_mode(cm_handle_t, int) { return CM_OK; }
int cm_set_extrinsic_tf(cm_handle_t, int, const double*, const double*) { return CM_OK; }
int cm_set_extrinsic(cm_handle_t, int, const float*, int) { return CM_OK; }
int cm_set_submit_policy(cm_handle_t h, int p) { std::lock_guard<std::mutex> lk(h->mu); h->policy = p; return CM_OK; }
int cm_submit_cloud(cm_handle_t h, int sensor, const void*, int64_t, const cm_layout_t*, uint64_t) {
  std::lock_guard<std::mutex> lk(h->mu);
  ++h->submits;
  if (h->policy == CM_SUBMIT_FIRST_WINS && ((h->submitted >> sensor) & 1ull)) { ++h->dropped; return CM_OK; }
  h->submitted |= 1ull << sensor;
  return CM_OK;
}
int cm_merge_frame_async(cm_handle_t h, uint64_t mask, int64_t* ticket) {
  std::lock_guard<std::mutex> lk(h->mu);
  if (!(h->submitted & mask)) return CM_E_NOT_READY;
  h->last_used = h->submitted & mask;
  h->submitted = 0;
  ++h->merges;
  *ticket = ++h->ticket;
  return CM_OK;
}
int cm_wait_frame(cm_handle_t h, int64_t, cm_frame_out_t* out, uint64_t* used, uint64_t* stamp) {
  std::lock_guard<std::mutex> lk(h->mu);
  if (out) { out->n_voxels = 0; out->n_survivors = 0; std::memset(&out->info, 0, sizeof(out->info)); }
  if (used) *used = h->last_used;
  if (stamp) *stamp = 0;
  return CM_OK;
}
// test hooks
long long cm_stub_counter(cm_handle_t h, int which) {
  std::lock_guard<std::mutex> lk(h->mu);
  return which == 0 ? h->submits : which == 1 ? h->dropped : which == 2 ? h->merges : (long long)h->last_used;
}
}
