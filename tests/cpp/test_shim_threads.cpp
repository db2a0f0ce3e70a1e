// test_shim_threads.cpp -- the hand-off between the six sensor-callback threads (ros::AsyncSpinner(6),
// pc_preprocessing_main.cpp:513) and the main loop (fusePointclouds, :131-160, :574) as FusedFrame implements it, run under
// ThreadSanitizer against a host-only stub of the C ABI (stub_cm.cpp). Checks: no data race (TSAN exits 66 on one), every
// fused frame holds all required sensors, and no delivery is lost: when every callback has delivered after a fusion, the
// next fuseAndVoxel succeeds.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "cloud_merger_shim.hpp"

extern "C" long long cm_stub_counter(cm_handle_t h, int which);
using namespace cloud_merger;

int main() {
  const int S = 6, ROUNDS = 3000;
  const uint64_t required = 0b101111;  // five required sensors, the top one (bit 4) optional (:134-136)
  int fail = 0;
  for (int first_wins = 0; first_wins < 2; ++first_wins) {
    FusedFrame ff(S, 64, required, 0, Params(), first_wins != 0);
    if (!ff.ok()) { std::printf("FAIL: FusedFrame\n"); return 1; }
    Cloud c;
    c.points.resize(8);
    c.width = 8; c.height = 1;
    std::atomic<bool> stop{false};
    std::vector<std::thread> th;
    for (int s = 0; s < S; ++s)
      th.emplace_back([&, s] {
        for (int i = 0; i < ROUNDS; ++i) {
          ff.onCloud(s, c);
          if ((i & 7) == s) std::this_thread::sleep_for(std::chrono::microseconds(20));  // sensors are not in lock step
          else std::this_thread::yield();
        }
      });
    long long fused = 0;
    std::thread main_loop([&] {
      Cloud f, v;
      while (!stop.load()) {
        if (ff.fuseAndVoxel(f, v)) {
          ++fused;
          const uint64_t used = (uint64_t)cm_stub_counter(ff.handle(), 3);
          if ((used & required) != required) { std::printf("FAIL: fused frame without a required sensor (%llx)\n", (unsigned long long)used); ++fail; }
        }
      }
    });
    for (auto& t : th) t.join();
    stop.store(true);
    main_loop.join();
    // no lost delivery: one more round of callbacks must open the gate again
    Cloud f, v;
    ff.fuseAndVoxel(f, v);  // drain whatever the last callbacks left
    for (int s = 0; s < S; ++s) ff.onCloud(s, c);
    if (!ff.ready() || !ff.fuseAndVoxel(f, v)) { std::printf("FAIL: the gate did not open after every sensor delivered\n"); ++fail; }
    if (ff.ready()) { std::printf("FAIL: flags not reset by the fusion\n"); ++fail; }
    const long long submits = cm_stub_counter(ff.handle(), 0), dropped = cm_stub_counter(ff.handle(), 1);
    std::printf("policy %s: %lld submissions (%lld dropped by the flag gate), %lld fused frames while the callbacks ran\n",
                first_wins ? "first-wins" : "latest-wins", submits, dropped, fused);
    if (submits != (long long)S * ROUNDS + S) { std::printf("FAIL: submissions lost\n"); ++fail; }
    if (!first_wins && dropped) { std::printf("FAIL: latest-wins must not drop\n"); ++fail; }
    if (fused < 10) { std::printf("FAIL: the main loop hardly ever got a frame (%lld)\n", fused); ++fail; }
  }
  std::printf("%s\n", fail ? "FAILED" : "shim threads ok");
  return fail ? 1 : 0;
}
