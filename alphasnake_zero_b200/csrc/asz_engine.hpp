// asz_engine.hpp -- host-side engine object behind the C ABI (include/asz_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <string>

#include "../../include/asz_b200.h"

namespace asz {

void set_error(const std::string& msg);
bool cuda_ok(cudaError_t err, const char* what);

#define ASZ_CUDA(call)                                         \
  do {                                                         \
    if (!::asz::cuda_ok((call), #call)) return ASZ_ERR_CUDA;   \
  } while (0)

struct SearchState;   // asz_mcts.cu
struct RecordStore;   // asz_records.cu

// Every entry point that takes an engine (or a network) runs on THAT object's device, whatever device is current in the
// calling thread ("one engine per GPU": several engines on different GPUs may live in one process).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess) prev = cur;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// NVTX range per phase (tic / encode, probe, net, sample, finish, records): visible in Nsight Systems / ncu --nvtx, a
// no-op function-pointer check when no tool is attached (tracing row)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
constexpr int kMaxDevices = 64;   // per-device "kernel attributes configured" flags (function attributes are per device)

// Device-resident game set: one record per game, structure-of-arrays over games.
//   cells  [n][PC]  u16 stamps (asz_common.cuh)
//   snakes [n][8]   u64 packed snake records
//   meta   [n][8]   u32 turn, episode, wall, body, head, starve, food_eaten, flags
struct GameSet {
  int n = 0;
  uint16_t* cells = nullptr;
  uint64_t* snakes = nullptr;
  uint32_t* meta = nullptr;
};

}  // namespace asz

struct asz_engine {
  asz_config cfg;
  int device = 0;
  int n_sm = 0;          // multiprocessors of `device`
  int pc = 0;            // padded cells per game
  int plane = 0;         // floats per plane
  int pitch = 0;         // floats between consecutive planes of the engine's own buffers (plane rounded up to 8: 32-byte rows)
  uint32_t chance_thresh = 0;
  asz::GameSet root;
  // env-step scratch owned by the engine
  float* planes = nullptr;        // [G*S][pitch]
  int32_t* row_ids = nullptr;     // [G*S]
  int32_t* row_count = nullptr;   // 8 MB of candidate locations for the kernel's one hot word (rows | game scheduler), see sched_off
  // Where inside that buffer the hot 64-bit word currently lives.  65,536 returning atomics per launch go to this one address, and
  // which L2 slice it is homed in decides whether the fused kernel can reach its fast regime at all: about half of the candidate
  // addresses leave it at ~240 us per launch whatever the state of the L2 (tools/env_bisect.py).  The physical
  // placement is not under the engine's control, so the L2 monitor rotates to the next candidate when sweeps do not help.
  size_t sched_off = 0;
  bool host_step = false;           // inside asz_env_step_host: results go to pinned host memory over PCIe, the launch is not judged
  bool hot_word_selected = false;   // asz_reset's probe has chosen sched_off for this process
  int32_t* rows_ptr() const { return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(row_count) + sched_off); }
  uint8_t* actions = nullptr;     // [G*8]
  int32_t* spawn_cells = nullptr; // [G]
  uint8_t* ended = nullptr;       // [G]
  int8_t* rewards = nullptr;      // [G*8]
  unsigned long long* totals = nullptr;  // [8]
  int device_hints = 1, host_hints = 0, step_hints = 1;   // L2 policy hints of env_step_kernel (asz_env.cu)
  // The fused tic + encode kernel of a large engine streams ~1 GB per launch through the L2 and is 1.5x slower when the L2 is full
  // of dirty lines (asz_env.cu, "L2 conditioning").  The engine knows when its own work has dirtied the L2 (bulk initialisation,
  // encode-only launches, the search and the network in between) and runs the read sweep before the next tic + encode launch.
  bool l2_dirty = true;
  int auto_condition = 1;          // ASZ_AUTO_CONDITION=0 disables (experiments)
  // ... and because a sweep does not always take (and other kernels of the application dirty the L2 behind the engine's back) it
  // MEASURES: every 16th streaming launch is bracketed by two events and its row count is copied to a pinned word; when the
  // sample shows the slow regime (plane bytes / time below l2_slow_gbs) the next launch is preceded by another sweep.
  struct L2Monitor {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int32_t* h_rows = nullptr;     // pinned
    bool pending = false, pending_host = false;
    int since_sample = 0, fails = 0, cooldown = 0;
    unsigned long long sweeps = 0, samples = 0, slow_samples = 0, rotations = 0;
    int candidate = 0;
    double last_gbs = 0.0;
  } l2mon;
  // asz_env_submit_host / asz_env_wait_host: two steps in flight, the inputs of the next one copied under the kernel of this one
  struct HostPipe {
    bool ready = false;
    cudaStream_t copy = nullptr;                 // host -> device copies of the steps' inputs
    cudaEvent_t copied[2] = {nullptr, nullptr};  // slot's inputs are on the device
    cudaEvent_t done[2] = {nullptr, nullptr};    // slot's launch and result copies have finished
    uint8_t* actions[2] = {nullptr, nullptr};    // [G*8] per slot
    int32_t* spawn[2] = {nullptr, nullptr};      // [G] per slot
    int32_t* h_rows = nullptr;                   // pinned, one 64-byte line per slot
    bool busy[2] = {false, false};
    int next = 0;
  } hostpipe;
  double l2_slow_gbs = 5500.0;     // ASZ_L2_SLOW_GBS: between the two regimes of a B200 (about 4,400 and 6,600 GB/s of plane bytes)
  asz::SearchState* search = nullptr;
  asz::RecordStore* records = nullptr;   // device-resident training records (asz_records_*)
};

namespace asz {
int gameset_alloc(GameSet& gs, int n, int pc);
void gameset_free(GameSet& gs);
int search_create(asz_engine* e);
void search_destroy(asz_engine* e);
void records_destroy(asz_engine* e);
void host_pipe_destroy(asz_engine* e);
}  // namespace asz
