// evaluator.cpp -- the evaluator front end between many CPU search threads and the device hot path
// (include/katacoffee_b200.h "Evaluator front end").
//
// Reference behaviour followed (cpp/neuralnet/):
//   NNEvaluator::evaluate     nneval.cpp:588-815   cache lookup, queue the row, wait, post-process, cache store
//   NNEvaluator::serve        nneval.cpp:341-586   a server thread takes whatever rows wait as soon as it is free
//   NNCacheTable              nneval.cpp:820-932   direct-mapped table, idx = hash0 & mask, striped mutex pool
//   NNInputs::getHash         nninputs.cpp:463-502 sit-hash (+ policy-temperature fold)
//   owner-map upgrade         nneval.cpp:612-623, 689-701   a hit without owner map is re-evaluated for the map only,
//                                                           policy and values of the cached entry are kept
//   owner-map post-processing nneval.cpp:817-838   tanh, flipped to white's perspective
//
// Built differently (nothing here is the reference's code):
//   * rows travel as packed positions (5 x u64) and the device does rules -> planes -> net -> masked softmax, reading the rows
//     from and writing the results to the mapped staging itself (no copy calls per batch);
//   * submit path: ticket = fetch_add(1); batch = ticket / maxBatch, slot = ticket % maxBatch; the client writes its row
//     into the staging of that batch and bumps its `ready` counter.  A server closes batch q by CAS-ing the ticket counter
//     up to (q+1)*maxBatch and waits for `ready` to reach the number of claimed slots;
//   * a staging buffer is recycled by the last client that has copied its result out (`consumed` == n);
//   * cache entries are inline rows of one float array.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <thread>

#ifndef KC_EVALUATOR_HOST_ONLY   // the thread-sanitizer build of tests/cpp/test_evaluator_stress.cpp compiles the queue and cache alone
#include "games.h"
#include "handle.h"
#include "net.h"
#else
#include "kc_internal.h"
#endif

namespace {

using kc::ZobristTables;

inline uint64_t splitMix64(uint64_t x) {  // cpp/core/hash.cpp:50-56
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
inline uint64_t basicLCong2(uint64_t x) { return 6364136223846793005ULL * x + 1442695040888963407ULL; }  // hash.cpp:32-35
constexpr uint64_t ZOBRIST_NN_POLICY_TEMP0 = 0xebcbdfeec6f4334bULL, ZOBRIST_NN_POLICY_TEMP1 = 0xb85e43ee243b5ad2ULL;  // nninputs.cpp:54

struct PackedRow {
  uint64_t black, white, hash0, hash1, misc;
  uint64_t blackHi = 0, whiteHi = 0;   // boards beyond 7x7: bits 64.. of the 128-bit bitboards (games_big.cuh)
};
// the criterion of kc_games_create: such boards run the 128-bit rules kernels, with their wider misc format
inline bool bigBoard(int W, int H) { return W > KC_MAX_DEVICE_LEN || H > KC_MAX_DEVICE_LEN; }

// The packing of kc_games_load (games.cu) for one position; also the literal nnHash and the cache key.
int packPosition(int W, int H, const kc_eval_position* p, float policyTemperature, PackedRow& r, uint64_t nnHash[2], uint64_t key[2]) {
  KC_CHECK(p && p->stones, "kc_evaluator: null position");
  KC_CHECK(p->nextPla == 1 || p->nextPla == 2, "kc_evaluator: nextPla must be 1 (black) or 2 (white)");
  KC_CHECK(p->numTurns >= 0 && p->numTurns <= 255, "kc_evaluator: numTurns out of range");
  const ZobristTables& z = kc::zobrist();
  const int HW = W * H, stride = W + 1;
  const bool big = bigBoard(W, H);   // 128-bit boards, 9-bit history entries, last direction at bit 45 (kc_games_load's wide branch)
  uint64_t bb = 0, ww = 0, bbHi = 0, wwHi = 0, a0 = z.sizeX[W][0] ^ z.sizeY[H][0], a1 = z.sizeX[W][1] ^ z.sizeY[H][1];
  for(int y = 0; y < H; y++)
    for(int x = 0; x < W; x++) {
      const int c = p->stones[y * W + x];
      KC_CHECK(c >= 0 && c <= 2, "kc_evaluator: stone colour must be 0, 1 or 2");
      if(c == 0) continue;
      const int idx = y * stride + x;
      const uint64_t bit = 1ULL << (idx & 63);
      if(idx < 64) { if(c == 1) bb |= bit; else ww |= bit; }
      else { if(c == 1) bbHi |= bit; else wwHi |= bit; }
      const int spot = (x + 1) + (y + 1) * (W + 1);   // board.h:74-75
      a0 ^= z.board[spot][c][0]; a1 ^= z.board[spot][c][1];
    }
  uint64_t m = 0;
  int lastDir = 4;
  if(p->moves) {
    for(int k = 0; k < 5; k++) {   // given oldest first; entry 0 of misc is the most recent
      const int pos = p->moves[(4 - k) * 2 + 0], pla = p->moves[(4 - k) * 2 + 1];
      if(pos < 0) continue;
      KC_CHECK(pos < 4 * HW && (pla == 1 || pla == 2), "kc_evaluator: bad history entry");
      if(big) m |= (uint64_t)((pos % HW) | (pla << 7)) << (9 * k);
      else m |= (uint64_t)((pos % HW) | (pla << 6)) << (8 * k);
      if(k == 0) lastDir = pos / HW;
    }
  }
  m |= ((uint64_t)lastDir << (big ? 45 : 40));
  // last five (cell, player) + last direction + min(numTurns, 5): with the stones, everything the planes and legality read -- the history
  // planes 7..10 are gated on numTurns >= 2..5 (games_device.cuh v1Planes, as nninputs.cpp:575-620), so two requests with equal moves
  // and different numTurns have different inputs and must not share a cache entry
  const uint64_t historyBits = m | ((uint64_t)std::min(p->numTurns, 5) << 44);
  m |= ((uint64_t)p->numTurns << 48) | ((uint64_t)(p->nextPla << 3) << 56);
  r.black = bb; r.white = ww; r.blackHi = bbHi; r.whiteHi = wwHi; r.hash0 = a0; r.hash1 = a1; r.misc = m;
  // NNInputs::getHash: getSitHash(nextPla) = pos_hash ^ ZOBRIST_PLAYER_HASH[pla] (board.cpp:288-292); never finished here
  uint64_t h0 = a0 ^ z.player[p->nextPla][0], h1 = a1 ^ z.player[p->nextPla][1];
  if(policyTemperature != 1.0f) {   // nninputs.cpp:485-492
    const int64_t t = (int64_t)(policyTemperature * 2048.0f);
    h0 ^= basicLCong2((uint64_t)t);
    h1 = splitMix64(h1 + (uint64_t)t);
    h0 += h1;
    h0 ^= ZOBRIST_NN_POLICY_TEMP0; h1 ^= ZOBRIST_NN_POLICY_TEMP1;
  }
  nnHash[0] = h0; nnHash[1] = h1;
  const uint64_t mix0 = splitMix64(historyBits ^ 0x6b4c1e2d9a5f3c71ULL), mix1 = splitMix64(mix0 ^ historyBits);
  key[0] = h0 ^ mix0; key[1] = h1 ^ mix1;
  return 0;
}

// ---- cache: direct-mapped, inline rows, striped mutexes (NNCacheTable's addressing, nneval.cpp:874-911) ------------------
struct Cache {
  uint64_t size = 0, mask = 0;
  uint32_t mutexMask = 0;
  int P = 0, HW = 0, rowFloats = 0;
  struct Tag { uint64_t k0, k1; uint8_t valid, hasOwner; int8_t symmetry; };
  std::unique_ptr<Tag[]> tags;
  std::unique_ptr<float[]> rows;   // [size][P + 4 + HW]: policy, win, loss, varTimeLeft, stError, owner map
  std::unique_ptr<std::mutex[]> mutexes;

  int init(int sizePow2, int mutexPow2, int P_, int HW_) {
    KC_CHECK(sizePow2 >= 0 && sizePow2 <= 30, "kc_evaluator: cacheSizePowerOfTwo out of range (0..30)");
    KC_CHECK(mutexPow2 >= 0 && mutexPow2 <= 24, "kc_evaluator: mutexPoolSizePowerOfTwo out of range (0..24)");
    if(mutexPow2 > sizePow2) mutexPow2 = sizePow2;   // nneval.cpp:860-861
    size = 1ULL << sizePow2; mask = size - 1; mutexMask = (1u << mutexPow2) - 1;
    P = P_; HW = HW_; rowFloats = P + 4 + HW;
    try {   // no C++ exception may cross the C ABI
      tags.reset(new Tag[size]());
      rows.reset(new float[size * (size_t)rowFloats]);
      mutexes.reset(new std::mutex[(size_t)mutexMask + 1]);
    } catch(const std::bad_alloc&) {
      return kc::fail("kc_evaluator: out of memory for a cache of 2^" + std::to_string(sizePow2) + " entries of " + std::to_string(rowFloats * 4 + 24) + " bytes");
    }
    return 0;
  }
  // true on a hit; *needOwner is set when the entry lacks the owner map the caller wants (policy / values are still copied)
  bool get(const uint64_t key[2], kc_eval_output* out, bool wantOwner, bool* needOwner) {
    const uint64_t idx = key[0] & mask;
    std::lock_guard<std::mutex> lock(mutexes[(uint32_t)idx & mutexMask]);
    const Tag& t = tags[idx];
    if(!t.valid || t.k0 != key[0] || t.k1 != key[1]) return false;
    const float* row = rows.get() + idx * (size_t)rowFloats;
    memcpy(out->policyProbs, row, (size_t)P * 4);
    out->whiteWinProb = row[P]; out->whiteLossProb = row[P + 1]; out->varTimeLeft = row[P + 2]; out->shorttermWinlossError = row[P + 3];
    out->symmetry = t.symmetry;
    *needOwner = wantOwner && !t.hasOwner;
    if(wantOwner && t.hasOwner) memcpy(out->whiteOwnerMap, row + P + 4, (size_t)HW * 4);
    return true;
  }
  void set(const uint64_t key[2], const kc_eval_output* out, bool hasOwner) {
    const uint64_t idx = key[0] & mask;
    std::lock_guard<std::mutex> lock(mutexes[(uint32_t)idx & mutexMask]);
    Tag& t = tags[idx];
    float* row = rows.get() + idx * (size_t)rowFloats;
    memcpy(row, out->policyProbs, (size_t)P * 4);
    row[P] = out->whiteWinProb; row[P + 1] = out->whiteLossProb; row[P + 2] = out->varTimeLeft; row[P + 3] = out->shorttermWinlossError;
    if(hasOwner) memcpy(row + P + 4, out->whiteOwnerMap, (size_t)HW * 4);
    t.k0 = key[0]; t.k1 = key[1]; t.valid = 1; t.hasOwner = hasOwner ? 1 : 0; t.symmetry = (int8_t)out->symmetry;
  }
  void clear() {
    for(uint64_t idx = 0; idx < size; idx++) {
      std::lock_guard<std::mutex> lock(mutexes[(uint32_t)idx & mutexMask]);
      tags[idx].valid = 0;
    }
  }
};

// ---- one staging buffer of the ring ------------------------------------------------------------------------------------
struct Staging {
  // inputs: the packed rows as five arrays (the layout of kc::State); mapped page-locked memory when the device backend is used
  uint64_t *black = nullptr, *white = nullptr, *hash0 = nullptr, *hash1 = nullptr, *misc = nullptr;
  uint64_t *blackHi = nullptr, *whiteHi = nullptr;   // boards beyond 7x7 only
  int8_t* symmetry = nullptr;
  uint8_t* wantOwner = nullptr;
  // outputs
  float *policy = nullptr, *winLoss = nullptr, *miscOut = nullptr, *ownership = nullptr;
  std::atomic<uint64_t> seq{0};     // the batch number this buffer currently serves
  std::atomic<int> ready{0};        // rows written by clients
  std::atomic<int> consumed{0};     // rows copied out by clients
  std::mutex m;
  std::condition_variable cv;       // completion of the batch, and recycling of the buffer
  uint64_t completedSeq = ~0ULL;    // under m
  std::atomic<uint64_t> completedFast{~0ULL};   // the same value, published after n / status: lets a client skip the mutex when its batch is already done
  int n = 0, status = 0;            // under m (written before completedSeq)
  std::string err;                  // under m
};

struct DeviceServer;   // below

}  // namespace

struct kc_evaluator {
  kc_evaluator_config cfg{};
  int W = 0, H = 0, HW = 0, P = 0;
  int ring = 0;
  bool pinned = false;
  std::unique_ptr<Staging[]> bufs;
  std::atomic<uint64_t> ticket{0};
  // servers
  std::mutex serverMutex;               // orders the servers: batches are closed in sequence
  std::condition_variable workCv;
  std::atomic<int> idleServers{0};
  uint64_t nextBatch = 0;               // under serverMutex
  bool killed = false;                  // under serverMutex
  std::vector<std::thread> threads;
  kc_eval_backend_fn fn = nullptr;
  void* user = nullptr;
  std::vector<DeviceServer*> devs;
  // cache + statistics
  std::unique_ptr<Cache> cache;
  std::atomic<uint64_t> rows{0}, batches{0}, hits{0}, misses{0}, upgrades{0}, backpressure{0};
};

namespace {

#ifndef KC_EVALUATOR_HOST_ONLY
// ---- the device backend of one server thread ---------------------------------------------------------------------------
// The staging buffers are page-locked and mapped (unified addressing: the device uses the host pointers), so a batch needs no
// copy calls: the rules kernel reads the packed rows straight from the staging, k_postprocess writes the results straight into it.
struct DeviceServer {
  kc_ctx* ctx = nullptr;
  kc_handle* handle = nullptr;
  kc_games* games = nullptr;        // maxBatch lanes; a batch of n rows runs with geom.numGames = n
  uint64_t* dH = nullptr;           // the device's own NNInputs::getHash of the rows (k_postprocess always writes it)
  int* dBad = nullptr;              // set by k_postprocess when a row's policy sum or win / loss probabilities are not finite
  int P = 0, HW = 0;

  int create(kc_ctx* c, const kc_model* model, const kc_evaluator_config& cfg) {
    ctx = c; HW = cfg.nnXLen * cfg.nnYLen; P = 4 * HW;
    if(kc_handle_create(c, model, cfg.maxBatch, cfg.nnXLen, cfg.nnYLen, cfg.handleFlags & ~KC_FLAG_INPUTS_NHWC, &handle)) return 1;
    if(kc_games_create(c, cfg.maxBatch, cfg.nnXLen, cfg.nnYLen, cfg.winLen, &games)) return 1;
    KC_CUDA(cudaMalloc(&dH, (size_t)cfg.maxBatch * 16));
    KC_CUDA(cudaMalloc(&dBad, 4));
    return 0;
  }
  void destroy() {
    if(ctx) cudaSetDevice(ctx->device);
    if(games) kc_games_destroy(games);
    if(handle) kc_handle_destroy(handle);
    games = nullptr; handle = nullptr;
    cudaFree(dH); cudaFree(dBad);
    dH = nullptr; dBad = nullptr;
  }
  int run(const kc_eval_batch* b) {
    KC_CUDA(cudaSetDevice(ctx->device));
    kc_games* G = games;
    KC_CHECK(b->n > 0 && b->n <= handle->maxBatch, "kc_evaluator: batch larger than maxBatch");
    const size_t n = (size_t)b->n;
    cudaStream_t st = G->stream;
    // the symmetries are read again by the trunk kernel's epilogue: those stay in device memory
    KC_CUDA(cudaMemcpyAsync(G->d_sym, b->symmetry, n, cudaMemcpyHostToDevice, st));
    const kc::State saved = G->st;
    const int savedN = G->geom.numGames;
    G->st.black = const_cast<uint64_t*>(b->black); G->st.white = const_cast<uint64_t*>(b->white);
    G->st.hash0 = const_cast<uint64_t*>(b->hash0); G->st.hash1 = const_cast<uint64_t*>(b->hash1);
    G->st.misc = const_cast<uint64_t*>(b->misc);   // (gameId stays the device array: the kernel loads it, nothing here depends on it)
    if(G->big) {
      KC_CHECK(b->blackHi && b->whiteHi, "kc_evaluator: a batch for a board beyond 7x7 without the upper bitboard words");
      G->st.blackHi = const_cast<uint64_t*>(b->blackHi); G->st.whiteHi = const_cast<uint64_t*>(b->whiteHi);
    }
    G->geom.numGames = b->n;
    const int rc = kc::gamesEval(G, handle, nullptr, nullptr, 0, false, /*symOnDevice=*/true);   // read-only on the state (no step)
    G->st = saved;
    G->geom.numGames = savedN;
    if(rc) return 1;
    KC_CUDA(cudaMemsetAsync(dBad, 0, 4, st));
    kc::launchPostprocess(handle, b->n, G->geom.LW, G->d_legal, G->d_status, G->d_sitHash, b->policyTemperature, b->policyProbs, b->whiteWinLoss,
                          b->miscOut, dH, st, 0, dBad);
    KC_CUDA(cudaGetLastError());
    if(b->wantOwnership) KC_CUDA(cudaMemcpyAsync(b->ownership, handle->d_own, n * HW * 4, cudaMemcpyDeviceToHost, st));
    int bad = 0;
    KC_CUDA(cudaMemcpyAsync(&bad, dBad, 4, cudaMemcpyDeviceToHost, st));
    KC_CUDA(cudaStreamSynchronize(st));
    if(kc::handleCheckAbort(handle)) return 1;
    // the reference throws here ("Got nonfinite for policy sum" / "Got nonfinite for nneval value", nneval.cpp:745-750, 789-793): the
    // batch fails -- its requests get the error, nothing of it enters the cache
    KC_CHECK(!bad, "kc_evaluator: got nonfinite for policy sum or nneval value (NaN / infinite weights or inputs?)");
    return 0;
  }
};
#else
struct DeviceServer {
  kc_ctx* ctx = nullptr;
  int run(const kc_eval_batch*) { return kc::fail("host-only build"); }
  void destroy() {}
};
#endif

void* stagingAlloc(bool pinned, size_t bytes) {
  void* p = nullptr;
#ifndef KC_EVALUATOR_HOST_ONLY
  if(pinned) {
    void* d = nullptr;
    if(cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    if(cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess || d != p) { cudaFreeHost(p); return nullptr; }   // kernels use the host pointers
  }
  else
#endif
    p = malloc(bytes);
  if(p) memset(p, 0, bytes);
  return p;
}
void stagingFree(bool pinned, void* p) {
  if(!p) return;
#ifndef KC_EVALUATOR_HOST_ONLY
  if(pinned) { cudaFreeHost(p); return; }
#endif
  free(p);
}

void serveLoop(kc_evaluator* ev, int serverIdx) {
  const uint64_t mb = (uint64_t)ev->cfg.maxBatch;
#ifndef KC_EVALUATOR_HOST_ONLY
  if(!ev->devs.empty()) cudaSetDevice(ev->devs[serverIdx]->ctx->device);
#endif
  for(;;) {
    uint64_t q; int n;
    {
      std::unique_lock<std::mutex> lock(ev->serverMutex);
      // sleep until a row has been claimed in the next batch (Dekker with the clients: they add to the ticket, then read idleServers)
      ev->idleServers.fetch_add(1);
      ev->workCv.wait(lock, [&] { return ev->killed || ev->ticket.load() > ev->nextBatch * mb; });
      ev->idleServers.fetch_sub(1);
      if(ev->killed) return;
      q = ev->nextBatch++;
      // close the batch: no ticket below (q+1)*mb may be handed out any more
      uint64_t t = ev->ticket.load();
      const uint64_t end = (q + 1) * mb;
      while(t < end && !ev->ticket.compare_exchange_weak(t, end)) {}
      n = (int)((t < end ? t : end) - q * mb);
    }
    Staging& b = ev->bufs[q & (uint64_t)(ev->ring - 1)];
    // the clients of this batch may still be packing their rows (or waiting for the buffer to be recycled)
    for(int spins = 0; b.seq.load(std::memory_order_acquire) != q || b.ready.load(std::memory_order_acquire) != n; spins++) {
      if(spins < 64) std::this_thread::yield();
      else std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    kc_eval_batch eb{};
    eb.n = n;
    eb.policyTemperature = ev->cfg.policyTemperature;
    eb.black = b.black; eb.white = b.white; eb.hash0 = b.hash0; eb.hash1 = b.hash1; eb.misc = b.misc;
    eb.blackHi = b.blackHi; eb.whiteHi = b.whiteHi;
    eb.symmetry = b.symmetry;
    eb.policyProbs = b.policy; eb.whiteWinLoss = b.winLoss; eb.miscOut = b.miscOut; eb.ownership = b.ownership;
    for(int i = 0; i < n; i++) if(b.wantOwner[i]) { eb.wantOwnership = 1; break; }
    std::string err;
    int status;
    if(ev->fn) {
      status = ev->fn(ev->user, serverIdx, &eb);
      if(status) err = "kc_evaluator: backend failed (status " + std::to_string(status) + ")";
    } else {
      status = ev->devs[serverIdx]->run(&eb);
      if(status) err = kc_last_error();
    }
    ev->rows.fetch_add((uint64_t)n, std::memory_order_relaxed);
    ev->batches.fetch_add(1, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lock(b.m);
      b.n = n; b.status = status; b.err = err;
      b.completedSeq = q;
      b.completedFast.store(q, std::memory_order_release);
    }
    b.cv.notify_all();
  }
}

struct Pending {   // a row in flight
  uint64_t q = 0; int slot = 0;
  uint64_t nnHash[2], key[2];
  bool cachedValues = false;   // owner-map upgrade: keep the cached policy / values
  int nextPla = 0;
};

bool bufferReady(Staging& b, uint64_t q) { return b.seq.load(std::memory_order_acquire) == q; }

void publishRow(kc_evaluator* ev, uint64_t t, const PackedRow& r, int symmetry, bool wantOwner, Pending& pd) {
  const uint64_t mb = (uint64_t)ev->cfg.maxBatch;
  const uint64_t q = t / mb; const int s = (int)(t % mb);
  Staging& b = ev->bufs[q & (uint64_t)(ev->ring - 1)];
  if(!bufferReady(b, q)) {
    ev->backpressure.fetch_add(1, std::memory_order_relaxed);
    std::unique_lock<std::mutex> lock(b.m);
    b.cv.wait(lock, [&] { return bufferReady(b, q); });
  }
  b.black[s] = r.black; b.white[s] = r.white; b.hash0[s] = r.hash0; b.hash1[s] = r.hash1; b.misc[s] = r.misc;
  if(b.blackHi) { b.blackHi[s] = r.blackHi; b.whiteHi[s] = r.whiteHi; }
  b.symmetry[s] = (int8_t)symmetry;
  b.wantOwner[s] = wantOwner ? 1 : 0;
  b.ready.fetch_add(1, std::memory_order_release);
  pd.q = q; pd.slot = s;
}

uint64_t claimTicket(kc_evaluator* ev) {
  const uint64_t t = ev->ticket.fetch_add(1);
  if(ev->idleServers.load() > 0) {
    std::lock_guard<std::mutex> lock(ev->serverMutex);
    ev->workCv.notify_one();
  }
  return t;
}

// wait for the batch, copy the row out, post-process the owner map, recycle the buffer when last
int collectRow(kc_evaluator* ev, const Pending& pd, bool wantOwner, kc_eval_output* out) {
  Staging& b = ev->bufs[pd.q & (uint64_t)(ev->ring - 1)];
  int n, status;
  std::string err;
  if(b.completedFast.load(std::memory_order_acquire) == pd.q && b.status == 0) {
    n = b.n; status = 0;   // n and status are stable until this row has been consumed
  } else {
    std::unique_lock<std::mutex> lock(b.m);
    b.cv.wait(lock, [&] { return b.completedSeq == pd.q; });
    n = b.n; status = b.status;
    if(status) err = b.err;
  }
  const int P = ev->P, HW = ev->HW, s = pd.slot;
  if(status == 0) {
    if(!pd.cachedValues) {
      memcpy(out->policyProbs, b.policy + (size_t)s * P, (size_t)P * 4);
      out->whiteWinProb = b.winLoss[2 * s]; out->whiteLossProb = b.winLoss[2 * s + 1];
      out->varTimeLeft = b.miscOut[2 * s]; out->shorttermWinlossError = b.miscOut[2 * s + 1];
      out->symmetry = b.symmetry[s];
    }
    if(wantOwner) {   // nneval.cpp:817-838
      const float* o = b.ownership + (size_t)s * HW;
      const float sign = pd.nextPla == 2 ? 1.f : -1.f;
      for(int i = 0; i < HW; i++) out->whiteOwnerMap[i] = sign * std::tanh(o[i]);
    }
  }
  if(b.consumed.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
    b.consumed.store(0, std::memory_order_relaxed);
    b.ready.store(0, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lock(b.m);
      b.seq.store(pd.q + (uint64_t)ev->ring, std::memory_order_release);
    }
    b.cv.notify_all();
  }
  if(status) return kc::fail(err);
  return 0;
}

int chooseSymmetry(const kc_evaluator* ev, int requested, const uint64_t key[2]) {
  if(requested != KC_SYMMETRY_NOTSPECIFIED) return requested;
  if(ev->cfg.doRandomize) return (int)(splitMix64(ev->cfg.randSeed ^ key[0]) >> 32) & 7;
  return ev->cfg.defaultSymmetry;
}

int checkRequest(const kc_evaluator* ev, const kc_eval_position* pos, int symmetry, int includeOwnerMap, const kc_eval_output* out) {
  KC_CHECK(ev && pos && out && out->policyProbs, "kc_evaluator_evaluate: null argument");
  KC_CHECK(symmetry >= -1 && symmetry <= 7, "kc_evaluator_evaluate: symmetry must be -1 (not specified) or 0..7");
  KC_CHECK(!includeOwnerMap || out->whiteOwnerMap, "kc_evaluator_evaluate: includeOwnerMap needs a whiteOwnerMap buffer");
  return 0;
}

// cache lookup + submit; returns 0 = submitted (pd valid), 1 = error, 2 = served from the cache
int beginRow(kc_evaluator* ev, const kc_eval_position* pos, int symmetry, bool skipCache, bool wantOwner, kc_eval_output* out, Pending& pd,
             PackedRow& row, int* symOut) {
  if(packPosition(ev->W, ev->H, pos, ev->cfg.policyTemperature, row, pd.nnHash, pd.key)) return 1;
  out->nnHash[0] = pd.nnHash[0]; out->nnHash[1] = pd.nnHash[1];
  out->cacheHit = 0;
  pd.nextPla = pos->nextPla;
  pd.cachedValues = false;
  if(ev->cache && !skipCache) {
    bool needOwner = false;
    if(ev->cache->get(pd.key, out, wantOwner, &needOwner)) {
      if(!needOwner) {
        ev->hits.fetch_add(1, std::memory_order_relaxed);
        out->cacheHit = 1;
        return 2;
      }
      pd.cachedValues = true;   // only the owner map is missing: nneval.cpp:612-623
      ev->upgrades.fetch_add(1, std::memory_order_relaxed);
    } else {
      ev->misses.fetch_add(1, std::memory_order_relaxed);
    }
  }
  *symOut = chooseSymmetry(ev, symmetry, pd.key);
  return 0;
}

}  // namespace

extern "C" {

static int evaluatorCreateCommon(const kc_evaluator_config* cfg, bool pinned, kc_evaluator** out) {
  KC_CHECK(cfg && out, "kc_evaluator_create: null argument");
  KC_CHECK(cfg->nnXLen >= 2 && cfg->nnYLen >= 2 && cfg->nnXLen <= KC_MAX_LEN && cfg->nnYLen <= KC_MAX_LEN,
           "kc_evaluator_create: board size must be within 2..10 (board.h:120)");
  KC_CHECK(cfg->winLen >= 2 && cfg->winLen <= 7, "kc_evaluator_create: winLen must be within 2..7");
  KC_CHECK(cfg->maxBatch > 0, "kc_evaluator_create: maxBatchSize is not positive");                // nneval.cpp:121-122
  KC_CHECK(cfg->maxConcurrentEvals > 0, "kc_evaluator_create: maxConcurrentEvals is not positive");  // nneval.cpp:119-120
  KC_CHECK(cfg->numServerThreads >= 1 && cfg->numServerThreads <= 64, "kc_evaluator_create: numServerThreads must be within 1..64");
  KC_CHECK(cfg->defaultSymmetry >= 0 && cfg->defaultSymmetry <= 7, "kc_evaluator_create: defaultSymmetry must be within 0..7");
  KC_CHECK(cfg->policyTemperature > 0.f, "kc_evaluator_create: policyTemperature must be positive");
  std::unique_ptr<kc_evaluator> ev(new kc_evaluator());
  ev->cfg = *cfg;
  ev->W = cfg->nnXLen; ev->H = cfg->nnYLen; ev->HW = ev->W * ev->H; ev->P = 4 * ev->HW;
  // nneval.cpp:128-136, plus one buffer per server: a batch stays in its staging buffer while it is processed (the reference's
  // servers swap the row list out of the ring instead)
  int r = cfg->maxConcurrentEvals / cfg->maxBatch + 3 + cfg->numServerThreads;
  int ring = 1;
  while(ring < r) ring *= 2;
  ev->ring = ring;
  ev->pinned = pinned;
  ev->bufs.reset(new Staging[ring]);
  const size_t mb = (size_t)cfg->maxBatch;
  for(int i = 0; i < ring; i++) {
    Staging& b = ev->bufs[i];
    b.seq.store((uint64_t)i);
    b.black = (uint64_t*)stagingAlloc(pinned, mb * 8); b.white = (uint64_t*)stagingAlloc(pinned, mb * 8);
    b.hash0 = (uint64_t*)stagingAlloc(pinned, mb * 8); b.hash1 = (uint64_t*)stagingAlloc(pinned, mb * 8);
    b.misc = (uint64_t*)stagingAlloc(pinned, mb * 8);
    if(bigBoard(ev->W, ev->H)) {
      b.blackHi = (uint64_t*)stagingAlloc(pinned, mb * 8); b.whiteHi = (uint64_t*)stagingAlloc(pinned, mb * 8);
      if(!(b.blackHi && b.whiteHi)) { kc_evaluator_destroy(ev.release()); return kc::fail("kc_evaluator_create: staging allocation failed"); }
    }
    b.symmetry = (int8_t*)stagingAlloc(pinned, mb);
    b.wantOwner = (uint8_t*)stagingAlloc(false, mb);
    b.policy = (float*)stagingAlloc(pinned, mb * ev->P * 4);
    b.winLoss = (float*)stagingAlloc(pinned, mb * 8); b.miscOut = (float*)stagingAlloc(pinned, mb * 8);
    b.ownership = (float*)stagingAlloc(pinned, mb * ev->HW * 4);
    if(!(b.black && b.white && b.hash0 && b.hash1 && b.misc && b.symmetry && b.wantOwner && b.policy && b.winLoss && b.miscOut && b.ownership)) {
      kc_evaluator_destroy(ev.release());
      return kc::fail("kc_evaluator_create: staging allocation failed");
    }
  }
  if(cfg->cacheSizePowerOfTwo >= 0) {
    ev->cache.reset(new Cache());
    if(ev->cache->init(cfg->cacheSizePowerOfTwo, cfg->mutexPoolSizePowerOfTwo, ev->P, ev->HW)) { kc_evaluator_destroy(ev.release()); return 1; }
  }
  *out = ev.release();
  return 0;
}

static int evaluatorSpawn(kc_evaluator* ev) {
  try {
    for(int i = 0; i < ev->cfg.numServerThreads; i++) ev->threads.emplace_back(serveLoop, ev, i);
  } catch(const std::exception& e) {
    return kc::fail(std::string("kc_evaluator_create: cannot start a server thread: ") + e.what());
  }
  return 0;
}

#ifndef KC_EVALUATOR_HOST_ONLY
int kc_evaluator_create_multi(int numServers, kc_ctx* const* ctxs, const kc_model* const* models, const kc_evaluator_config* cfg, kc_evaluator** out) {
  KC_CHECK(ctxs && models && cfg && out, "kc_evaluator_create: null argument");
  KC_CHECK(numServers == cfg->numServerThreads, "kc_evaluator_create_multi: one (context, model) pair per server thread is required");
  for(int i = 0; i < numServers; i++) {
    KC_CHECK(ctxs[i] && models[i], "kc_evaluator_create: null context or model");
    KC_CHECK(models[i]->ctx == ctxs[i], "kc_evaluator_create: a model must have been created on its server's context");
  }
  KC_CUDA(cudaSetDevice(ctxs[0]->device));
  kc_evaluator* ev = nullptr;
  if(evaluatorCreateCommon(cfg, true, &ev)) return 1;
  for(int i = 0; i < numServers; i++) {
    DeviceServer* d = new DeviceServer();
    ev->devs.push_back(d);
    if(d->create(ctxs[i], models[i], *cfg)) {
      const std::string msg = kc_last_error();
      kc_evaluator_destroy(ev);
      return kc::fail(msg);
    }
  }
  if(evaluatorSpawn(ev)) { const std::string msg = kc_last_error(); kc_evaluator_destroy(ev); return kc::fail(msg); }
  *out = ev;
  return 0;
}

int kc_evaluator_create(kc_ctx* ctx, const kc_model* model, const kc_evaluator_config* cfg, kc_evaluator** out) {
  KC_CHECK(ctx && model && cfg, "kc_evaluator_create: null argument");
  KC_CHECK(cfg->numServerThreads >= 1 && cfg->numServerThreads <= 64, "kc_evaluator_create: numServerThreads must be within 1..64");
  std::vector<kc_ctx*> ctxs((size_t)cfg->numServerThreads, ctx);
  std::vector<const kc_model*> models((size_t)cfg->numServerThreads, model);
  return kc_evaluator_create_multi(cfg->numServerThreads, ctxs.data(), models.data(), cfg, out);
}
#endif

int kc_evaluator_create_custom(const kc_evaluator_config* cfg, kc_eval_backend_fn fn, void* user, kc_evaluator** out) {
  KC_CHECK(fn, "kc_evaluator_create_custom: null backend function");
  kc_evaluator* ev = nullptr;
  if(evaluatorCreateCommon(cfg, false, &ev)) return 1;
  ev->fn = fn; ev->user = user;
  if(evaluatorSpawn(ev)) { const std::string msg = kc_last_error(); kc_evaluator_destroy(ev); return kc::fail(msg); }
  *out = ev;
  return 0;
}

int kc_evaluator_destroy(kc_evaluator* ev) {
  if(!ev) return 0;
  {
    std::lock_guard<std::mutex> lock(ev->serverMutex);
    ev->killed = true;
  }
  ev->workCv.notify_all();
  for(std::thread& t : ev->threads) t.join();
  for(DeviceServer* d : ev->devs) { d->destroy(); delete d; }
  if(ev->bufs)
    for(int i = 0; i < ev->ring; i++) {
      Staging& b = ev->bufs[i];
      const bool p = ev->pinned;
      stagingFree(p, b.blackHi); stagingFree(p, b.whiteHi);
      stagingFree(p, b.black); stagingFree(p, b.white); stagingFree(p, b.hash0); stagingFree(p, b.hash1); stagingFree(p, b.misc);
      stagingFree(p, b.symmetry); stagingFree(false, b.wantOwner);
      stagingFree(p, b.policy); stagingFree(p, b.winLoss); stagingFree(p, b.miscOut); stagingFree(p, b.ownership);
    }
  delete ev;
  return 0;
}

int kc_evaluator_evaluate(kc_evaluator* ev, const kc_eval_position* pos, int symmetry, int skipCache, int includeOwnerMap,
                          kc_eval_output* out) {
  if(checkRequest(ev, pos, symmetry, includeOwnerMap, out)) return 1;
  Pending pd; PackedRow row; int sym = 0;
  const int st = beginRow(ev, pos, symmetry, skipCache != 0, includeOwnerMap != 0, out, pd, row, &sym);
  if(st == 1) return 1;
  if(st == 2) return 0;
  publishRow(ev, claimTicket(ev), row, sym, includeOwnerMap != 0, pd);
  if(collectRow(ev, pd, includeOwnerMap != 0, out)) return 1;
  if(ev->cache) ev->cache->set(pd.key, out, includeOwnerMap != 0);   // nneval.cpp:841-843 (also when skipCache)
  return 0;
}

int kc_evaluator_evaluate_many(kc_evaluator* ev, int n, const kc_eval_position* pos, const int8_t* symmetryOrNull, int skipCache,
                               int includeOwnerMap, kc_eval_output* out) {
  KC_CHECK(ev && n >= 0 && (n == 0 || (pos && out)), "kc_evaluator_evaluate_many: bad argument");
  for(int i = 0; i < n; i++)
    if(checkRequest(ev, pos + i, symmetryOrNull ? symmetryOrNull[i] : KC_SYMMETRY_NOTSPECIFIED, includeOwnerMap, out + i)) return 1;
  const bool wantOwner = includeOwnerMap != 0;
  const uint64_t mb = (uint64_t)ev->cfg.maxBatch;
  std::vector<Pending> pend((size_t)n);
  std::vector<int> inFlight;   // indices submitted and not yet collected
  inFlight.reserve((size_t)n);
  int failed = 0;
  auto drain = [&] {
    for(int j : inFlight) {
      if(collectRow(ev, pend[j], wantOwner, out + j)) failed = 1;
      else if(ev->cache) ev->cache->set(pend[j].key, out + j, wantOwner);
    }
    inFlight.clear();
  };
  for(int i = 0; i < n; i++) {
    PackedRow row; int sym = 0;
    const int st = beginRow(ev, pos + i, symmetryOrNull ? symmetryOrNull[i] : KC_SYMMETRY_NOTSPECIFIED, skipCache != 0, wantOwner, out + i, pend[i], row, &sym);
    if(st == 1) { drain(); return 1; }
    if(st == 2) continue;
    const uint64_t t = claimTicket(ev);
    // this thread must not wait for a buffer while it holds results that keep an older buffer from being recycled
    if(!inFlight.empty() && !bufferReady(ev->bufs[(t / mb) & (uint64_t)(ev->ring - 1)], t / mb)) drain();
    publishRow(ev, t, row, sym, wantOwner, pend[i]);
    inFlight.push_back(i);
  }
  drain();
  return failed ? kc::fail("kc_evaluator: the backend failed on a batch") : 0;
}

int kc_evaluator_clear_cache(kc_evaluator* ev) {
  KC_CHECK(ev, "kc_evaluator_clear_cache: null evaluator");
  if(ev->cache) ev->cache->clear();
  return 0;
}

int kc_evaluator_get_stats(const kc_evaluator* ev, kc_evaluator_stats* out) {
  KC_CHECK(ev && out, "kc_evaluator_get_stats: null argument");
  out->rowsProcessed = ev->rows.load(); out->batchesProcessed = ev->batches.load();
  out->cacheHits = ev->hits.load(); out->cacheMisses = ev->misses.load(); out->ownerMapUpgrades = ev->upgrades.load();
  out->backpressureWaits = ev->backpressure.load();
  return 0;
}

int kc_evaluator_clear_stats(kc_evaluator* ev) {
  KC_CHECK(ev, "kc_evaluator_clear_stats: null evaluator");
  ev->rows = 0; ev->batches = 0; ev->hits = 0; ev->misses = 0; ev->upgrades = 0; ev->backpressure = 0;
  return 0;
}

int kc_eval_position_hash(int xSize, int ySize, const kc_eval_position* pos, float policyTemperature, uint64_t nnHash[2], uint64_t cacheKey[2]) {
  KC_CHECK(xSize >= 2 && ySize >= 2 && xSize <= KC_MAX_LEN && ySize <= KC_MAX_LEN, "kc_eval_position_hash: board size out of range");
  KC_CHECK(nnHash && cacheKey, "kc_eval_position_hash: null argument");
  PackedRow r;
  return packPosition(xSize, ySize, pos, policyTemperature, r, nnHash, cacheKey);
}

int kc_eval_unpack_position(int xSize, int ySize, uint64_t black, uint64_t white, uint64_t misc, int8_t* stones, int8_t* nextPla,
                            int16_t* movesCellPla, int32_t* numTurns, int32_t* lastDir) {
  KC_CHECK(xSize >= 2 && ySize >= 2 && xSize <= KC_MAX_DEVICE_LEN && ySize <= KC_MAX_DEVICE_LEN,
           "kc_eval_unpack_position: board size out of range (beyond 7x7: kc_eval_unpack_position_wide)");
  KC_CHECK(stones, "kc_eval_unpack_position: null argument");
  const int stride = xSize + 1;
  for(int y = 0; y < ySize; y++)
    for(int x = 0; x < xSize; x++) {
      const uint64_t bit = 1ULL << (y * stride + x);
      stones[y * xSize + x] = (black & bit) ? 1 : (white & bit) ? 2 : 0;
    }
  if(nextPla) *nextPla = (int8_t)((misc >> 59) & 3);
  if(numTurns) *numTurns = (int32_t)((misc >> 48) & 0xff);
  if(lastDir) *lastDir = (int32_t)((misc >> 40) & 0xff);
  if(movesCellPla)
    for(int k = 0; k < 5; k++) {   // oldest first, like kc_eval_position::moves; (cell, player), cell = -1 for none
      const int byte = (int)((misc >> (8 * k)) & 0xff), pla = byte >> 6;
      movesCellPla[(4 - k) * 2 + 0] = pla ? (int16_t)(byte & 63) : (int16_t)-1;
      movesCellPla[(4 - k) * 2 + 1] = (int16_t)pla;
    }
  return 0;
}

int kc_eval_unpack_position_wide(int xSize, int ySize, uint64_t blackLo, uint64_t blackHi, uint64_t whiteLo, uint64_t whiteHi, uint64_t misc,
                                 int8_t* stones, int8_t* nextPla, int16_t* movesCellPla, int32_t* numTurns, int32_t* lastDir) {
  KC_CHECK(xSize >= 2 && ySize >= 2 && xSize <= KC_MAX_LEN && ySize <= KC_MAX_LEN && bigBoard(xSize, ySize),
           "kc_eval_unpack_position_wide: for boards beyond 7x7, up to 10x10");
  KC_CHECK(stones, "kc_eval_unpack_position_wide: null argument");
  const int stride = xSize + 1;
  for(int y = 0; y < ySize; y++)
    for(int x = 0; x < xSize; x++) {
      const int idx = y * stride + x;
      const uint64_t bit = 1ULL << (idx & 63);
      const bool b = ((idx < 64 ? blackLo : blackHi) & bit) != 0, w = ((idx < 64 ? whiteLo : whiteHi) & bit) != 0;
      stones[y * xSize + x] = b ? 1 : w ? 2 : 0;
    }
  if(nextPla) *nextPla = (int8_t)((misc >> 59) & 3);
  if(numTurns) *numTurns = (int32_t)((misc >> 48) & 0xff);
  if(lastDir) *lastDir = (int32_t)((misc >> 45) & 7);
  if(movesCellPla)
    for(int k = 0; k < 5; k++) {   // oldest first, like kc_eval_position::moves; (cell, player), cell = -1 for none
      const int e = (int)((misc >> (9 * k)) & 0x1ff), pla = e >> 7;
      movesCellPla[(4 - k) * 2 + 0] = pla ? (int16_t)(e & 127) : (int16_t)-1;
      movesCellPla[(4 - k) * 2 + 1] = (int16_t)pla;
    }
  return 0;
}

}  // extern "C"
