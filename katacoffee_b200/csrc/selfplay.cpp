// kc_selfplay_run: self-play over several GPUs of one box from ONE process -- the reference's sharding, natively.
// The reference runs one NN server thread + one ComputeHandle per GPU (cpp/program/setup.cpp:167-229: gpuIdxByServerThread) under
// numGameThreads game loops (cpp/command/selfplay.cpp:390-392) that hand finished games to a data-writer thread
// (selfplay.cpp:226-260, TrainingDataWriter::writeGame / flushIfNonempty, cpp/dataio/trainingwrite.cpp:779-1062).
// Here every GPU gets one host thread that owns a context, the weights, a compute handle and a search pool of G games
// (csrc/search.cu: the whole game loop is on the device), plus one writer thread that turns the finished games' rows into the
// reference's .npz files while the next chunk of moves is being searched.  Games never cross devices: the only exchange is one
// ncclReduce(sum) of the statistics counters at the end (NCCL is loaded with dlopen, so the library has no link-time dependency on it).
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "kc_internal.h"

namespace {

constexpr uint64_t GAME_ID_STRIDE = 1ULL << 40;   // ids of pool i live in [first + i * 2^40, ...): never collide, refills included
constexpr int NUM_COUNTERS = 12;                  // kc_search_stats as uint64[12]

struct RowBuffers {
  std::vector<uint8_t> bin; std::vector<float> glob; std::vector<int16_t> policy; std::vector<float> targets; std::vector<int8_t> value;
  int rows = 0;
  void size(int maxRows, int HW) {
    bin.resize((size_t)maxRows * 15 * ((HW + 7) / 8)); glob.resize(maxRows); policy.resize((size_t)maxRows * 2 * 4 * HW);
    targets.resize((size_t)maxRows * 64); value.resize((size_t)maxRows * 5 * HW);
  }
};

// one writer thread per pool: the device thread fills buffer k and posts it; the writer deflates it into a file while the device
// thread searches the next chunk (two buffers, so the device thread waits only if a file takes longer than a chunk of moves)
struct Writer {
  RowBuffers buf[2];
  int posted = 0, written = 0;         // chunks posted / finished
  bool stop = false, failed = false;
  std::string error, dir;
  int pool = 0, W = 0, H = 0;
  uint64_t files = 0, rows = 0, bytes = 0;
  std::mutex m;
  std::condition_variable cv;
  std::thread th;
  void start() {
    th = std::thread([this] {
      for(;;) {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return stop || written < posted; });
        if(written >= posted) return;
        const int k = written;
        lk.unlock();
        RowBuffers& b = buf[k & 1];
        if(b.rows > 0 && !dir.empty()) {
          char name[64];
          snprintf(name, sizeof name, "/pool%02d_%06d.npz", pool, k);
          const std::string path = dir + name;
          if(kc_training_write_npz(path.c_str(), b.rows, W, H, b.bin.data(), b.glob.data(), b.policy.data(), b.targets.data(), b.value.data())) {
            lk.lock(); failed = true; error = kc_last_error(); written++; cv.notify_all(); continue;
          }
          if(FILE* f = fopen(path.c_str(), "rb")) { fseek(f, 0, SEEK_END); bytes += (uint64_t)ftell(f); fclose(f); }
          files++;
        }
        rows += (uint64_t)b.rows;
        lk.lock();
        written++;
        cv.notify_all();
      }
    });
  }
  RowBuffers& acquire(int k) {   // buffer for chunk k: free once chunk k - 2 has been written
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return written >= k - 1; });
    return buf[k & 1];
  }
  void post() { { std::lock_guard<std::mutex> lk(m); posted++; } cv.notify_all(); }
  void finish() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv.notify_all();
    if(th.joinable()) th.join();
  }
};

struct Pool {
  int index = 0, device = 0;
  kc_ctx* ctx = nullptr; kc_model* model = nullptr; kc_handle* handle = nullptr; kc_search* search = nullptr;
  kc_search_stats stats{};
  double deviceMs = 0;
  uint64_t dropped = 0;
  int64_t launches = 0;
  Writer writer;
  std::string error;
  void release() {
    if(search) kc_search_destroy(search);
    if(handle) kc_handle_destroy(handle);
    if(model) kc_model_destroy(model);
    if(ctx) kc_ctx_destroy(ctx);
    search = nullptr; handle = nullptr; model = nullptr; ctx = nullptr;
  }
};

// sense-reversing barrier over the pool threads (C++17: no std::barrier)
struct Barrier {
  std::mutex m; std::condition_variable cv; int n, waiting = 0, phase = 0;
  explicit Barrier(int count) : n(count) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const int ph = phase;
    if(++waiting == n) { waiting = 0; phase++; cv.notify_all(); }
    else cv.wait(lk, [&] { return phase != ph; });
  }
};

// ---- NCCL through dlopen: ncclCommInitAll / ncclReduce / ncclGroupStart / ncclGroupEnd / ncclCommDestroy (nccl.h) ----
typedef struct ncclComm* ncclComm_t;
struct Nccl {
  void* lib = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  bool load() {
    for(const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if(lib) break;
    }
    if(!lib) return false;
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
    Reduce = reinterpret_cast<decltype(Reduce)>(dlsym(lib, "ncclReduce"));
    return CommInitAll && CommDestroy && GroupStart && GroupEnd && Reduce;
  }
};
constexpr int NCCL_UINT64 = 5, NCCL_SUM = 0;   // ncclDataType_t ncclUint64, ncclRedOp_t ncclSum (nccl.h, stable since NCCL 2.0)

// sum of every pool's counters on the first device, over NVLink; false (and `why`) if NCCL is not usable
bool reduceWithNccl(const std::vector<Pool>& pools, uint64_t total[NUM_COUNTERS], std::string& why) {
  Nccl nccl;
  if(!nccl.load()) { why = "libnccl.so.2 not loadable"; return false; }
  const int n = (int)pools.size();
  std::vector<int> devs(n);
  for(int i = 0; i < n; i++) devs[i] = pools[i].device;
  std::vector<ncclComm_t> comms(n, nullptr);
  if(nccl.CommInitAll(comms.data(), n, devs.data()) != 0) { why = "ncclCommInitAll failed"; return false; }
  std::vector<uint64_t*> dbuf(n, nullptr);
  std::vector<cudaStream_t> streams(n, nullptr);
  bool ok = true;
  for(int i = 0; i < n && ok; i++) {
    ok = cudaSetDevice(devs[i]) == cudaSuccess && cudaMalloc(&dbuf[i], 2 * NUM_COUNTERS * 8) == cudaSuccess && cudaStreamCreate(&streams[i]) == cudaSuccess &&
         cudaMemcpyAsync(dbuf[i], &pools[i].stats, NUM_COUNTERS * 8, cudaMemcpyHostToDevice, streams[i]) == cudaSuccess;
  }
  if(ok) {
    nccl.GroupStart();
    for(int i = 0; i < n; i++) {
      cudaSetDevice(devs[i]);
      if(nccl.Reduce(dbuf[i], dbuf[i] + NUM_COUNTERS, NUM_COUNTERS, NCCL_UINT64, NCCL_SUM, 0, comms[i], streams[i]) != 0) ok = false;
    }
    if(nccl.GroupEnd() != 0) ok = false;
  }
  if(ok) {
    cudaSetDevice(devs[0]);
    ok = cudaMemcpyAsync(total, dbuf[0] + NUM_COUNTERS, NUM_COUNTERS * 8, cudaMemcpyDeviceToHost, streams[0]) == cudaSuccess;
    for(int i = 0; i < n; i++) { cudaSetDevice(devs[i]); if(cudaStreamSynchronize(streams[i]) != cudaSuccess) ok = false; }
  }
  for(int i = 0; i < n; i++) {
    cudaSetDevice(devs[i]);
    if(dbuf[i]) cudaFree(dbuf[i]);
    if(streams[i]) cudaStreamDestroy(streams[i]);
    if(comms[i]) nccl.CommDestroy(comms[i]);
  }
  if(!ok) why = "ncclReduce failed";
  return ok;
}

}  // namespace

extern "C" int kc_selfplay_run(const kc_selfplay_config* cfg, const kc_model_desc* desc, const kc_search_params* params, kc_search_stats* total,
                               kc_selfplay_report* report) {
  static_assert(sizeof(kc_search_stats) == NUM_COUNTERS * 8, "kc_search_stats is NUM_COUNTERS uint64 counters");
  KC_CHECK(cfg && desc && params, "kc_selfplay_run: null argument");
  KC_CHECK(cfg->numDevices > 0 && cfg->numDevices <= 64 && cfg->devices, "kc_selfplay_run: need 1..64 devices");
  KC_CHECK(cfg->gamesPerDevice > 0 && cfg->moves > 0, "kc_selfplay_run: gamesPerDevice and moves must be positive");
  KC_CHECK(cfg->warmupMoves >= 0 && cfg->staggerPlies >= 0 && cfg->maxRowsPerChunk >= 0, "kc_selfplay_run: negative count");
  const int n = cfg->numDevices, G = cfg->gamesPerDevice, W = cfg->xSize, H = cfg->ySize, HW = W * H;
  const int chunkMoves = cfg->movesPerChunk > 0 ? cfg->movesPerChunk : cfg->moves;
  const bool wantRows = cfg->maxRowsPerChunk > 0;
  std::vector<Pool> pools(n);
  Barrier ready(n + 1), done(n + 1);
  std::atomic<int> failed{0};
  std::vector<std::thread> threads;
  for(int i = 0; i < n; i++) {
    Pool& P = pools[i];
    P.index = i; P.device = cfg->devices[i];
    threads.emplace_back([&, i] {
      Pool& P = pools[i];
      auto bad = [&](int status, const char* what) {
        if(status == 0) return false;
        P.error = std::string(what) + ": " + kc_last_error();
        failed.store(1);
        return true;
      };
      // ---- set-up (untimed): context, weights, handle, search pool, staggered starts, warm-up moves
      bool ok = !bad(kc_ctx_create(P.device, &P.ctx), "kc_ctx_create") && !bad(kc_model_create(P.ctx, desc, &P.model), "kc_model_create") &&
                !bad(kc_handle_create(P.ctx, P.model, G, W, H, cfg->handleFlags, &P.handle), "kc_handle_create") &&
                !bad(kc_search_create(P.ctx, P.handle, G, W, H, cfg->winLen, params, &P.search), "kc_search_create") &&
                !bad(kc_search_reset(P.search, cfg->seed, cfg->firstGameId + (uint64_t)i * GAME_ID_STRIDE), "kc_search_reset");
      if(ok && cfg->staggerPlies > 0) {
        // the steady-state mix of a running self-play: lane g starts after (g mod staggerPlies) random-legal plies, as games that
        // were refilled at different times do (movePos -2 = the counter-RNG move, -1 = stay)
        std::vector<int16_t> mv(G);
        for(int t = 0; t < cfg->staggerPlies && ok; t++) {
          for(int g = 0; g < G; g++) mv[g] = (g % cfg->staggerPlies) > t ? -2 : -1;
          ok = !bad(kc_games_step(kc_search_games(P.search), mv.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "kc_games_step");
        }
      }
      if(ok && wantRows) ok = !bad(kc_search_enable_training_rows(P.search, cfg->maxRowsPerChunk), "kc_search_enable_training_rows");
      if(ok && cfg->warmupMoves > 0) ok = !bad(kc_search_play(P.search, cfg->warmupMoves, nullptr, nullptr, nullptr), "kc_search_play (warm-up)");
      if(ok && wantRows) {
        int r = 0, d = 0;   // rows of the warm-up are discarded
        ok = !bad(kc_search_read_training_rows(P.search, &r, &d, nullptr, nullptr, nullptr, nullptr, nullptr, 1), "kc_search_read_training_rows");
        P.writer.pool = i; P.writer.W = W; P.writer.H = H; P.writer.dir = cfg->outputDir ? cfg->outputDir : "";
        for(RowBuffers& b : P.writer.buf) b.size(cfg->maxRowsPerChunk, HW);
        P.writer.start();
      }
      const int64_t l0 = ok ? kc_search_launch_count(P.search) : 0;
      ready.wait();
      // ---- timed part: chunks of moves; after each the finished games' rows go to the writer thread
      int chunk = 0;
      for(int m = 0; ok && m < cfg->moves && !failed.load(); m += chunkMoves, chunk++) {
        float ms = 0.f;
        ok = !bad(kc_search_play(P.search, std::min(chunkMoves, cfg->moves - m), nullptr, &P.stats, &ms), "kc_search_play");
        P.deviceMs += ms;
        if(ok && wantRows) {
          RowBuffers& b = P.writer.acquire(chunk);
          int d = 0;
          ok = !bad(kc_search_read_training_rows(P.search, &b.rows, &d, b.bin.data(), b.glob.data(), b.policy.data(), b.targets.data(), b.value.data(), 1),
                    "kc_search_read_training_rows");
          P.dropped += (uint64_t)d;
          if(ok) P.writer.post();
        }
      }
      if(wantRows) {
        P.writer.finish();
        if(P.writer.failed && ok) { P.error = "kc_training_write_npz: " + P.writer.error; failed.store(1); }
      }
      if(P.search) P.launches = kc_search_launch_count(P.search) - l0;
      done.wait();
    });
  }
  ready.wait();
  const auto t0 = std::chrono::steady_clock::now();
  done.wait();
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for(std::thread& t : threads) t.join();
  std::string err;
  for(Pool& P : pools) if(!P.error.empty() && err.empty()) err = "pool " + std::to_string(P.index) + " (device " + std::to_string(P.device) + "): " + P.error;
  // ---- the one exchange of the whole job: sum of the counters on the first device
  uint64_t sum[NUM_COUNTERS] = {0};
  int usedNccl = 0;
  std::string why;
  if(err.empty()) {
    if(n > 1 && !cfg->noNccl && reduceWithNccl(pools, sum, why)) usedNccl = 1;
    else
      for(const Pool& P : pools) { const uint64_t* s = reinterpret_cast<const uint64_t*>(&P.stats); for(int k = 0; k < NUM_COUNTERS; k++) sum[k] += s[k]; }
  }
  if(report) {
    memset(report, 0, sizeof *report);
    report->wallSeconds = wall; report->reducedWithNccl = usedNccl;
    for(const Pool& P : pools) {
      report->deviceMsMax = std::max(report->deviceMsMax, P.deviceMs);
      report->rowsWritten += P.writer.rows; report->rowsDropped += P.dropped; report->filesWritten += P.writer.files; report->bytesWritten += P.writer.bytes;
      report->kernelLaunches += (uint64_t)P.launches;
    }
  }
  if(total) memcpy(total, sum, sizeof sum);
  for(Pool& P : pools) { cudaSetDevice(P.device); P.release(); }
  if(!err.empty()) return kc::fail("kc_selfplay_run: " + err);
  return 0;
}
