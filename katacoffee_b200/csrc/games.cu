// Batched Coffee rules, history, sit-hash and NNInputs V1 feature planes on the device.
//
// One thread owns one game for the rule logic (pure 64-bit bitboard arithmetic, ~600 SASS
// instructions per ply; one warp = 32 games per CTA), then the whole CTA expands the 15 plane bitboards of its games into the
// output tensor with 128-bit coalesced stores -- the plane write is the HBM-bound part (1500 B per
// position in fp32), the logic is noise next to it.
//
// Semantics (reference file:line; canonical readings from SURVEY.md 8.1 quoted as "ledger X"):
//   Board::isLegal                 cpp/game/board.cpp:185-227   (ledger B)
//   Board::playMoveAssumeLegal     cpp/game/board.cpp:427-435
//   Board::maxConsecutives/checkGameEnd  board.cpp:315-335,376-383 (ledger N: overlines win)
//   BoardHistory::makeBoardMoveAssumeLegal  cpp/game/boardhistory.cpp:157-176 (ledger C draw, D history)
//   Board::getSitHash              cpp/game/board.cpp:288-292   (ledger E: literal, no lastLoc)
//   NNInputs::fillRowV1            cpp/neuralnet/nninputs.cpp:508-657 (ledger F, G)
//   copyInputsWithSymmetry         cpp/neuralnet/nninputs.cpp:252-357 (ledger K parity mode)
//   random-legal playout           cpp/program/playutils.cpp:10-32 with the counter RNG of SURVEY 8(d)
//
// Bitboard layout: bit = y*(W+1) + x, i.e. one always-zero pad column per row, so that the four
// line directions are plain shifts (N-S: W+1, W-E: 1, NW-SE: W+2, NE-SW: W) that cannot wrap.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kc_internal.h"
#include "net.h"

#include "games_device.cuh"
#include "games_big.cuh"

namespace kc {

// CTA shapes.  FEAT 0/1/2 (no planes, fp32 NCHW, fp32 NHWC): 64 threads = 64 games, every thread runs the rules of
// its own game; the fp32 NCHW fast path is warp-autonomous (a warp expands the 32 games it stepped, no CTA barrier),
// so warps drift apart and the plane stores of one overlap the bitboard arithmetic of another.
// FEAT 3 (bf16 trunk tiles): 128 threads, NB*8 <= 32 games = whole trunk tiles per CTA.
constexpr int TB_PLAIN = 64;
constexpr int TB_TILES = 128;

// ---------------------------------------------------------------------------------------------
// The kernel.  DO_STEP: play one ply first.  FEAT: 0 none, 1 fp32 NCHW, 2 fp32 NHWC, 3 bf16 trunk tiles.
// ---------------------------------------------------------------------------------------------
struct FeatOut {
  float* planes;            // FEAT 1/2
  float* global;            // FEAT 1/2
  uint4* tiles;             // FEAT 3: [tile][2][128] 16-byte chunks
  const int8_t* symmetry;   // per game or null
  int permuteDirs;          // play mode (ledger K): with a symmetry, channels 3..6 follow symDir
  int gamesPerBlock;
  uint32_t tileOne, tileK;  // FEAT 3: bit patterns of 1.0 and of win_len in the handle's 16-bit operand format (fp16 or bf16: both exact)
};

template <bool DO_STEP, int FEAT, class D>
__global__ void __launch_bounds__(FEAT >= 3 ? TB_TILES : TB_PLAIN) games_kernel(const Geom g, State st, const int16_t* __restrict__ moves,
                                                        int useMoves, const uint64_t* __restrict__ zob,
                                                        StepOut so, FeatOut fo) {
  constexpr int THREADS = FEAT >= 3 ? TB_TILES : TB_PLAIN;
  // most games a CTA can hold.  FEAT 4 = FEAT 3 (bf16 trunk tiles) with one work item (two tiles) per CTA: 1.5 KB of shared memory, so the
  // kernel fits beside a resident trunk kernel (the search's half batches)
  constexpr int GPB = FEAT == 3 ? 32 : FEAT == 4 ? 8 : TB_PLAIN;
  __shared__ uint64_t sPlanes[15][GPB + 1];
  __shared__ uint8_t sSrcPad[8][52];   // [symmetry][dst cell] -> padded bit index of the source cell
  __shared__ int8_t sSym[GPB];
  // FEAT 1 without symmetry: a warp's whole output as one bit string (bit e = output float e of its 32 games), so the
  // expansion is one shared load + one 128-bit store per four floats
  __shared__ uint32_t sBits[FEAT == 1 ? GPB * 15 * 49 / 32 + 4 : 1];
  const bool fastNCHW = (FEAT == 1) && fo.symmetry == nullptr;
  const D dm(g);
  using BB = typename D::BB;
  const int gpb = fo.gamesPerBlock;
  const int gBase = blockIdx.x * gpb;
  const int t = threadIdx.x;
  const int gi = gBase + t;
  const bool active = t < gpb && gi < g.numGames;

  if(FEAT == 1 && fastNCHW) {
    // each warp clears the bit string of its own 32 games (32*E bits = E words, word-aligned)
    const int E = 15 * dm.HW();
    for(int i = (t & 31); i < E; i += 32) sBits[(t >> 5) * E + i] = 0;
    __syncwarp();
  } else if(FEAT != 0) {
    // symmetry tables: dst[sym(h,w)] = src[h,w] (nninputs.cpp:252-335), built per CTA (<= 8*49 entries)
    for(int i = t; i < 8 * dm.HW(); i += THREADS) {
      int sym = i / dm.HW(), cell = i % dm.HW();
      int h = cell / dm.W(), w = cell % dm.W();
      bool tr = (sym & 4) && dm.H() == dm.W(), fx = (sym & 2) != 0, fy = (sym & 1) != 0;
      if(tr) { bool tmp = fx; fx = fy; fy = tmp; }
      int rowStep = dm.W(), colStep = 1, base = 0;
      if(fy) { base += (dm.H() - 1) * rowStep; rowStep = -rowStep; }
      if(fx) { base += (dm.W() - 1) * colStep; colStep = -colStep; }
      if(tr) { int tmp = rowStep; rowStep = colStep; colStep = tmp; }
      int dst = base + h * rowStep + w * colStep;
      sSrcPad[sym][dst] = (uint8_t)(cell + cell / dm.W());
    }
    if(t < gpb) sSym[t] = (active && fo.symmetry) ? fo.symmetry[gi] : 0;
  }

  unsigned long long cSteps = 0, cFin = 0, cB = 0, cW = 0, cD = 0, cXor = 0;
  if(active) {
    GameRegs<BB> s;
    s.black = (BB)st.black[gi]; s.white = (BB)st.white[gi]; s.h0 = st.hash0[gi]; s.h1 = st.hash1[gi];
    s.id = st.gameId[gi]; s.misc = st.misc[gi];
    BB L[4];
    int played = -1;
    bool illegal = false;
    if(DO_STEP) {
      int mv = useMoves ? (int)moves[gi] : -1;
      played = stepGame(dm, g, s, mv, useMoves != 0, zob, L, illegal);
      st.black[gi] = (uint64_t)s.black; st.white[gi] = (uint64_t)s.white; st.hash0[gi] = s.h0; st.hash1[gi] = s.h1;
      st.gameId[gi] = s.id; st.misc[gi] = s.misc;
    } else {
      BB empty = (BB)g.all & ~(s.black | s.white);
      int lastCell = histPla(s.misc, 0) ? histCell(s.misc, 0) : -1;
      legalMasks(dm, g, empty, lastCell, lastDirOf(s.misc), L);
    }
    int fl = flagsOf(s.misc);
    int nextPla = (fl >> 3) & 3;
    uint64_t sh0 = s.h0 ^ g.playerHash[nextPla][0], sh1 = s.h1 ^ g.playerHash[nextPla][1];
    {
      if(so.status) so.status[gi] = (uint32_t)numTurnsOf(s.misc) | ((uint32_t)(fl & 1) << 8) | ((uint32_t)((fl >> 1) & 3) << 9) |
                                    ((uint32_t)nextPla << 11) | (illegal ? (1u << 15) : 0u);
      if(so.sitHash) { so.sitHash[2 * (size_t)gi] = sh0; so.sitHash[2 * (size_t)gi + 1] = sh1; }
      if(so.played) so.played[gi] = (int16_t)played;
      if(so.legal) {
        // policy order: bit = dir*HW + y*W + x (nninputs.cpp:6-14)
        uint64_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for(int d = 0; d < 4; d++) {
          uint64_t dense = toDense(dm, L[d]);
          int off = d * dm.HW(), w = off >> 6, sh = off & 63;
#pragma unroll
          for(int q = 0; q < 4; q++) {
            if(q == w) acc[q] |= dense << sh;
            if(q == w + 1 && sh) acc[q] |= dense >> (64 - sh);
          }
        }
        for(int wd = 0; wd < g.LW; wd++)
          so.legal[(size_t)gi * g.LW + wd] = (uint32_t)(acc[wd >> 1] >> ((wd & 1) * 32));
      }
      if(DO_STEP && played >= 0) {
        cSteps = 1;
        cXor = sh0;
        if(fl & 1) { cFin = 1; int wn = (fl >> 1) & 3; cB = wn == 1; cW = wn == 2; cD = wn == 0; }
      }
    }
    if(FEAT != 0) {
      if(fastNCHW) {
        v1Planes(dm, g, s, L, &sPlanes[0][t], GPB + 1);
        const int E = 15 * dm.HW();
        uint32_t bitPos = (uint32_t)t * E, wi = bitPos >> 5;
        int fill = bitPos & 31;
        uint64_t acc = 0;
        auto append = [&](uint64_t bits, int n) {
          acc |= bits << fill;
          fill += n;
          if(fill >= 32) { atomicOr(&sBits[wi++], (uint32_t)acc); acc >>= 32; fill -= 32; }
        };
#pragma unroll 1
        for(int c = 0; c < 15; c++) {
          uint64_t D = toDense(dm, (BB)sPlanes[c][t]);
          if(dm.HW() <= 25) append(D, dm.HW());
          else { append(D & 0x1FFFFFFULL, 25); append(D >> 25, dm.HW() - 25); }
        }
        if(fill) atomicOr(&sBits[wi], (uint32_t)acc);
      } else {
        v1Planes(dm, g, s, L, &sPlanes[0][t], GPB + 1);
      }
    }
  } else if(FEAT != 0 && t < GPB) {
#pragma unroll
    for(int c = 0; c < 15; c++) sPlanes[c][t] = 0;
  }

  if(DO_STEP && so.stats) {
    // warp-reduce then one atomic per warp per counter
    for(int o = 16; o > 0; o >>= 1) {
      cSteps += __shfl_xor_sync(0xffffffffu, cSteps, o);
      cFin += __shfl_xor_sync(0xffffffffu, cFin, o);
      cB += __shfl_xor_sync(0xffffffffu, cB, o);
      cW += __shfl_xor_sync(0xffffffffu, cW, o);
      cD += __shfl_xor_sync(0xffffffffu, cD, o);
      cXor ^= __shfl_xor_sync(0xffffffffu, cXor, o);
    }
    if((t & 31) == 0) {
      if(cSteps) atomicAdd(&so.stats[0], cSteps);
      if(cFin) atomicAdd(&so.stats[2], cFin);
      if(cB) atomicAdd(&so.stats[3], cB);
      if(cW) atomicAdd(&so.stats[4], cW);
      if(cD) atomicAdd(&so.stats[5], cD);
      if(cXor) atomicXor(&so.stats[6], cXor);
    }
  }
  if(FEAT == 0) return;
  const int ng = min(gpb, g.numGames - gBase);
  if(FEAT == 1 && fastNCHW) {
    // warp-autonomous expansion: no CTA barrier, the warp's own 32 games only
    __syncwarp();
    const int w = t >> 5, lane = t & 31;
    const int ngw = min(32, ng - w * 32);
    if(ngw <= 0) return;
    const int E = 15 * dm.HW();
    const int total = ngw * E;
    const uint32_t* bits = sBits + w * E;
    float* out = fo.planes + ((size_t)gBase + w * 32) * E;
    float4* out4 = reinterpret_cast<float4*>(out);
    const int nvec = total >> 2;
#pragma unroll 4
    for(int j = lane; j < nvec; j += 32) {
      uint32_t b = bits[j >> 3] >> ((j & 7) * 4);
      float4 v = make_float4((float)(b & 1u), (float)((b >> 1) & 1u), (float)((b >> 2) & 1u), (float)((b >> 3) & 1u));
      __stcs(&out4[j], v);
    }
    for(int e = (nvec << 2) + lane; e < total; e += 32) out[e] = (float)((bits[e >> 5] >> (e & 31)) & 1u);
    if(fo.global && lane < ngw) fo.global[gBase + w * 32 + lane] = (float)dm.K();
    return;
  }
  __syncthreads();
  if(ng <= 0) return;
  if(false) {
  } else if(FEAT == 1 || FEAT == 2) {
    const int E = 15 * dm.HW();
    const uint32_t magicE = (uint32_t)((0x100000000ULL + E - 1) / E);
    const int inner = FEAT == 1 ? dm.HW() : 15;
    const uint32_t magicI = (uint32_t)((0x100000000ULL + inner - 1) / inner);
    const int total = ng * E;
    float* out = fo.planes + (size_t)gBase * E;
    auto elem = [&](int e) -> float {
      int gl = (int)__umulhi((uint32_t)e, magicE);
      int r = e - gl * E;
      int q = (int)__umulhi((uint32_t)r, magicI);
      int rem = r - q * inner;
      int c = FEAT == 1 ? q : rem;
      int cell = FEAT == 1 ? rem : q;
      int bit = sSrcPad[sSym[gl]][cell];
      if(fo.permuteDirs) c = playModeChannel(c, sSym[gl]);
      return (float)((sPlanes[c][gl] >> bit) & 1ULL);
    };
    const int nvec = total >> 2;
    float4* out4 = reinterpret_cast<float4*>(out);
    for(int j = t; j < nvec; j += THREADS) {
      int e = j << 2;
      float4 v = make_float4(elem(e), elem(e + 1), elem(e + 2), elem(e + 3));
      __stcs(&out4[j], v);
    }
    for(int e = (nvec << 2) + t; e < total; e += THREADS) out[e] = elem(e);
    if(fo.global && t < ng) fo.global[gBase + t] = (float)dm.K();   // nninputs.cpp:656
  } else if(FEAT >= 3) {
    // trunk input tiles: row = y*tileRowW + b*(W+1) + x, 16 bf16 channels per row as two 16 B chunks
    // (channels 0..14 = V1 planes, channel 15 = the global feature win_len broadcast on board cells)
    const int ntiles = (ng + g.NB - 1) / g.NB;
    const int tileBase = gBase / g.NB;
    const uint32_t ONE = fo.tileOne;
    const uint32_t kval = fo.tileK;
    for(int j = t; j < ntiles * 256; j += THREADS) {
      int tile = j >> 8, chunk = (j >> 7) & 1, row = j & 127;
      int y = row / g.tileRowW, rr = row - y * g.tileRowW;
      int b = rr / dm.stride(), x = rr - b * dm.stride();
      int gl = tile * g.NB + b;
      uint4 v = make_uint4(0, 0, 0, 0);
      if(y < dm.H() && x < dm.W() && gl < ng) {
        int bit = sSrcPad[sSym[gl]][y * dm.W() + x];
        uint32_t w[4];
#pragma unroll
        for(int q = 0; q < 4; q++) {
          int c0 = chunk * 8 + 2 * q, c1 = c0 + 1;
          if(fo.permuteDirs) { c0 = playModeChannel(c0, sSym[gl]); if(c1 < 15) c1 = playModeChannel(c1, sSym[gl]); }
          uint32_t lo = ((sPlanes[c0][gl] >> bit) & 1ULL) ? ONE : 0u;
          uint32_t hi = (c1 < 15) ? (((sPlanes[c1][gl] >> bit) & 1ULL) ? ONE : 0u) : kval;
          w[q] = lo | (hi << 16);
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fo.tiles[((size_t)(tileBase + tile) * 2 + chunk) * 128 + row] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Several plies in one launch (rules + fp32 NCHW planes, random-legal moves, no symmetry): the state of a game stays
// in registers between plies, and because every warp is autonomous the bitboard arithmetic of one warp's next ply
// overlaps the plane stores of the others -- the single-ply kernel pays a store-free rules phase on every launch.
// Ply p writes its planes AND its masks, status words, hashes and moves to ring slot (firstSlot + p) % 4, so the outputs of the
// last four plies of a launch are all in HBM for a consumer (kc_games_read_run_ply); the host picks firstSlot so that the last
// ply lands in slot 0 = the object's ordinary buffers.  Semantically identical to `plies` launches of games_kernel<true, 1>.
// ---------------------------------------------------------------------------------------------
struct PlaneRing { float* slot[4]; };

template <class D>
__global__ void __launch_bounds__(TB_PLAIN) games_multi_kernel(const Geom g, State st, const uint64_t* __restrict__ zob, StepOut so,
                                                               PlaneRing ring, float* __restrict__ global, int plies, int firstSlot) {
  constexpr int GPB = TB_PLAIN;
  __shared__ uint64_t sPlanes[15][GPB + 1];
  __shared__ uint32_t sBits[GPB * 15 * 49 / 32 + 4];
  const D dm(g);
  using BB = typename D::BB;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int gBase = blockIdx.x * GPB;
  const int gi = gBase + t;
  const bool active = gi < g.numGames;
  const int E = 15 * dm.HW();
  const int ngw = min(32, g.numGames - gBase - w * 32);
  if(ngw <= 0) return;
  GameRegs<BB> s;
  s.black = 0; s.white = 0; s.h0 = s.h1 = s.id = s.misc = 0;
  if(active) {
    s.black = (BB)st.black[gi]; s.white = (BB)st.white[gi]; s.h0 = st.hash0[gi]; s.h1 = st.hash1[gi];
    s.id = st.gameId[gi]; s.misc = st.misc[gi];
  }
  unsigned long long cSteps = 0, cFin = 0, cB = 0, cW = 0, cD = 0, cXor = 0;
  uint32_t* bits = sBits + w * E;
  for(int p = 0; p < plies; p++) {
    for(int i = lane; i < E; i += 32) bits[i] = 0;
    __syncwarp();
    if(active) {
      const size_t gs = (size_t)((firstSlot + p) & 3) * g.numGames + gi;   // this ply's slot of the per-ply output ring
      BB L[4];
      bool illegal = false;
      const int played = stepGame(dm, g, s, -1, false, zob, L, illegal);
      const int fl = flagsOf(s.misc);
      const int nextPla = (fl >> 3) & 3;
      const uint64_t sh0 = s.h0 ^ g.playerHash[nextPla][0], sh1 = s.h1 ^ g.playerHash[nextPla][1];
      if(so.status) so.status[gs] = (uint32_t)numTurnsOf(s.misc) | ((uint32_t)(fl & 1) << 8) | ((uint32_t)((fl >> 1) & 3) << 9) | ((uint32_t)nextPla << 11);
      if(so.sitHash) { so.sitHash[2 * gs] = sh0; so.sitHash[2 * gs + 1] = sh1; }
      if(so.played) so.played[gs] = (int16_t)played;
      if(so.legal) {
        uint64_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for(int d = 0; d < 4; d++) {
          uint64_t dense = toDense(dm, L[d]);
          int off = d * dm.HW(), wq = off >> 6, sh = off & 63;
#pragma unroll
          for(int q = 0; q < 4; q++) {
            if(q == wq) acc[q] |= dense << sh;
            if(q == wq + 1 && sh) acc[q] |= dense >> (64 - sh);
          }
        }
        for(int wd = 0; wd < g.LW; wd++) so.legal[gs * g.LW + wd] = (uint32_t)(acc[wd >> 1] >> ((wd & 1) * 32));
      }
      if(played >= 0) {
        cSteps += 1;
        cXor ^= sh0;
        if(fl & 1) { cFin += 1; int wn = (fl >> 1) & 3; cB += wn == 1; cW += wn == 2; cD += wn == 0; }
      }
      v1Planes(dm, g, s, L, &sPlanes[0][t], GPB + 1);
      uint32_t bitPos = (uint32_t)t * E, wi = bitPos >> 5;
      int fill = bitPos & 31;
      uint64_t acc = 0;
      auto append = [&](uint64_t b, int n) {
        acc |= b << fill;
        fill += n;
        if(fill >= 32) { atomicOr(&sBits[wi++], (uint32_t)acc); acc >>= 32; fill -= 32; }
      };
#pragma unroll 1
      for(int c = 0; c < 15; c++) {
        uint64_t Dn = toDense(dm, (BB)sPlanes[c][t]);
        if(dm.HW() <= 25) append(Dn, dm.HW());
        else { append(Dn & 0x1FFFFFFULL, 25); append(Dn >> 25, dm.HW() - 25); }
      }
      if(fill) atomicOr(&sBits[wi], (uint32_t)acc);
    }
    __syncwarp();
    const int total = ngw * E;
    float* out = ring.slot[(firstSlot + p) & 3] + ((size_t)gBase + w * 32) * E;
    float4* out4 = reinterpret_cast<float4*>(out);
    const int nvec = total >> 2;
#pragma unroll 4
    for(int j = lane; j < nvec; j += 32) {
      uint32_t b = bits[j >> 3] >> ((j & 7) * 4);
      float4 v = make_float4((float)(b & 1u), (float)((b >> 1) & 1u), (float)((b >> 2) & 1u), (float)((b >> 3) & 1u));
      __stcs(&out4[j], v);
    }
    for(int e = (nvec << 2) + lane; e < total; e += 32) out[e] = (float)((bits[e >> 5] >> (e & 31)) & 1u);
    __syncwarp();
  }
  if(global && lane < ngw) global[gBase + w * 32 + lane] = (float)dm.K();
  if(active) {
    st.black[gi] = (uint64_t)s.black; st.white[gi] = (uint64_t)s.white; st.hash0[gi] = s.h0; st.hash1[gi] = s.h1;
    st.gameId[gi] = s.id; st.misc[gi] = s.misc;
  }
  if(so.stats) {
    for(int o = 16; o > 0; o >>= 1) {
      cSteps += __shfl_xor_sync(0xffffffffu, cSteps, o);
      cFin += __shfl_xor_sync(0xffffffffu, cFin, o);
      cB += __shfl_xor_sync(0xffffffffu, cB, o);
      cW += __shfl_xor_sync(0xffffffffu, cW, o);
      cD += __shfl_xor_sync(0xffffffffu, cD, o);
      cXor ^= __shfl_xor_sync(0xffffffffu, cXor, o);
    }
    if(lane == 0) {
      if(cSteps) atomicAdd(&so.stats[0], cSteps);
      if(cFin) atomicAdd(&so.stats[2], cFin);
      if(cB) atomicAdd(&so.stats[3], cB);
      if(cW) atomicAdd(&so.stats[4], cW);
      if(cD) atomicAdd(&so.stats[5], cD);
      if(cXor) atomicXor(&so.stats[6], cXor);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same several-plies launch with the two phases of a ply on different warps: a CTA of 128 threads owns 64 games; warps 0, 1
// (producers) step their 32 games and pack the planes into a shared bit string, warps 2, 3 (consumers) expand the bit string
// of the producer two warps below them into fp32 and stream it out.  Bit strings are double-buffered and handed over with
// named barriers (bar.arrive / bar.sync on 64 threads: full[pair][buf], empty[pair][buf]), so a producer runs the bitboard
// arithmetic of ply p+1 while its consumer is still storing ply p.  The grid is still one CTA per 64 games, but twice the warps
// are resident (28 instead of 14 per SM at 65,536 games), and the store stream no longer pauses for the rules phase.
// Same outputs as games_multi_kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <class D, int NC>   // NC consumer warps per producer warp
__global__ void __launch_bounds__((1 + NC) * TB_PLAIN) games_multi_split_kernel(const Geom g, State st, const uint64_t* __restrict__ zob, StepOut so,
                                                                         PlaneRing ring, float* __restrict__ global, int plies, int firstSlot) {
  constexpr int GPB = TB_PLAIN;                       // games per CTA
  constexpr int WORDS = 15 * 49;                      // bit-string words of 32 games (E bits each), sized for 7x7
  __shared__ uint64_t sPlanes[15][GPB + 1];
  __shared__ uint32_t sBits[2][2][WORDS + 1];         // [pair][buffer]
  const D dm(g);
  using BB = typename D::BB;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int pair = w & 1, role = w >> 1;              // role 0: producer (rules + pack), 1..NC: consumers (expand + store)
  constexpr int NTHR = 32 * (1 + NC);                 // threads meeting at a hand-over barrier
  const int gBase = blockIdx.x * GPB;
  const int E = 15 * dm.HW();
  const int ngw = min(32, g.numGames - gBase - pair * 32);
  if(ngw <= 0) return;                                // both warps of the pair leave together
  const int barFull = 1 + pair * 4, barEmpty = 3 + pair * 4;   // + buffer index
  if(role >= 1) {
    const int ci = role - 1;
    for(int p = 0; p < plies; p++) {
      const int buf = p & 1;
      named_sync(barFull + buf, NTHR);
      const uint32_t* bits = sBits[pair][buf];
      const int total = ngw * E;
      float* out = ring.slot[(firstSlot + p) & 3] + ((size_t)gBase + pair * 32) * E;
      float4* out4 = reinterpret_cast<float4*>(out);
      const int nvec = total >> 2;
#pragma unroll 4
      for(int j = lane + 32 * ci; j < nvec; j += 32 * NC) {
        uint32_t b = bits[j >> 3] >> ((j & 7) * 4);
        float4 v = make_float4((float)(b & 1u), (float)((b >> 1) & 1u), (float)((b >> 2) & 1u), (float)((b >> 3) & 1u));
        __stcs(&out4[j], v);
      }
      if(ci == 0) for(int e = (nvec << 2) + lane; e < total; e += 32) out[e] = (float)((bits[e >> 5] >> (e & 31)) & 1u);
      if(p + 2 < plies) named_arrive(barEmpty + buf, NTHR);   // the producer may refill this buffer
    }
    if(ci == 0 && global && lane < ngw) global[gBase + pair * 32 + lane] = (float)dm.K();
    return;
  }
  // ---- producer
  const int tl = pair * 32 + lane;                    // game within the CTA
  const int gi = gBase + tl;
  const bool active = gi < g.numGames;
  GameRegs<BB> s;
  s.black = 0; s.white = 0; s.h0 = s.h1 = s.id = s.misc = 0;
  if(active) {
    s.black = (BB)st.black[gi]; s.white = (BB)st.white[gi]; s.h0 = st.hash0[gi]; s.h1 = st.hash1[gi];
    s.id = st.gameId[gi]; s.misc = st.misc[gi];
  }
  unsigned long long cSteps = 0, cFin = 0, cB = 0, cW = 0, cD = 0, cXor = 0;
  for(int p = 0; p < plies; p++) {
    const int buf = p & 1;
    if(p >= 2) named_sync(barEmpty + buf, NTHR);      // the consumers have finished with this buffer (ply p - 2)
    uint32_t* bits = sBits[pair][buf];
    for(int i = lane; i < E; i += 32) bits[i] = 0;
    __syncwarp();
    if(active) {
      const size_t gs = (size_t)((firstSlot + p) & 3) * g.numGames + gi;   // this ply's slot of the per-ply output ring
      BB L[4];
      bool illegal = false;
      const int played = stepGame(dm, g, s, -1, false, zob, L, illegal);
      const int fl = flagsOf(s.misc);
      const int nextPla = (fl >> 3) & 3;
      const uint64_t sh0 = s.h0 ^ g.playerHash[nextPla][0], sh1 = s.h1 ^ g.playerHash[nextPla][1];
      if(so.status) so.status[gs] = (uint32_t)numTurnsOf(s.misc) | ((uint32_t)(fl & 1) << 8) | ((uint32_t)((fl >> 1) & 3) << 9) | ((uint32_t)nextPla << 11);
      if(so.sitHash) { so.sitHash[2 * gs] = sh0; so.sitHash[2 * gs + 1] = sh1; }
      if(so.played) so.played[gs] = (int16_t)played;
      if(so.legal) {
        uint64_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for(int d = 0; d < 4; d++) {
          uint64_t dense = toDense(dm, L[d]);
          int off = d * dm.HW(), wq = off >> 6, sh = off & 63;
#pragma unroll
          for(int q = 0; q < 4; q++) {
            if(q == wq) acc[q] |= dense << sh;
            if(q == wq + 1 && sh) acc[q] |= dense >> (64 - sh);
          }
        }
        for(int wd = 0; wd < g.LW; wd++) so.legal[gs * g.LW + wd] = (uint32_t)(acc[wd >> 1] >> ((wd & 1) * 32));
      }
      if(played >= 0) {
        cSteps += 1;
        cXor ^= sh0;
        if(fl & 1) { cFin += 1; int wn = (fl >> 1) & 3; cB += wn == 1; cW += wn == 2; cD += wn == 0; }
      }
      v1Planes(dm, g, s, L, &sPlanes[0][tl], GPB + 1);
      uint32_t bitPos = (uint32_t)lane * E, wi = bitPos >> 5;
      int fill = bitPos & 31;
      uint64_t acc = 0;
      auto append = [&](uint64_t b, int n) {
        acc |= b << fill;
        fill += n;
        if(fill >= 32) { atomicOr(&bits[wi++], (uint32_t)acc); acc >>= 32; fill -= 32; }
      };
#pragma unroll 1
      for(int c = 0; c < 15; c++) {
        uint64_t Dn = toDense(dm, (BB)sPlanes[c][tl]);
        if(dm.HW() <= 25) append(Dn, dm.HW());
        else { append(Dn & 0x1FFFFFFULL, 25); append(Dn >> 25, dm.HW() - 25); }
      }
      if(fill) atomicOr(&bits[wi], (uint32_t)acc);
    }
    __syncwarp();
    named_arrive(barFull + buf, NTHR);                // hand the bit string of this ply to the consumers
  }
  if(active) {
    st.black[gi] = (uint64_t)s.black; st.white[gi] = (uint64_t)s.white; st.hash0[gi] = s.h0; st.hash1[gi] = s.h1;
    st.gameId[gi] = s.id; st.misc[gi] = s.misc;
  }
  if(so.stats) {
    for(int o = 16; o > 0; o >>= 1) {
      cSteps += __shfl_xor_sync(0xffffffffu, cSteps, o);
      cFin += __shfl_xor_sync(0xffffffffu, cFin, o);
      cB += __shfl_xor_sync(0xffffffffu, cB, o);
      cW += __shfl_xor_sync(0xffffffffu, cW, o);
      cD += __shfl_xor_sync(0xffffffffu, cD, o);
      cXor ^= __shfl_xor_sync(0xffffffffu, cXor, o);
    }
    if(lane == 0) {
      if(cSteps) atomicAdd(&so.stats[0], cSteps);
      if(cFin) atomicAdd(&so.stats[2], cFin);
      if(cB) atomicAdd(&so.stats[3], cB);
      if(cW) atomicAdd(&so.stats[4], cW);
      if(cD) atomicAdd(&so.stats[5], cD);
      if(cXor) atomicXor(&so.stats[6], cXor);
    }
  }
}

}  // namespace kc

// =============================================================================================
// Host side
// =============================================================================================
#include "games.h"

namespace {
using namespace kc;

// BASELINE configs[4] (6x6 k=4) also has a compile-time instantiation (64-bit bitboards, unrolled); KC_GAMES_STATIC6=0 disables it
bool static6() { static const bool on = [] { const char* e = getenv("KC_GAMES_STATIC6"); return !e || atoi(e) != 0; }(); return on; }

template <bool DO_STEP, class D>
void launchGamesD(kc_games* G, int feat, int useMoves, const StepOut& so, const FeatOut& fo, int blocks) {
  switch(feat) {
    case 0: games_kernel<DO_STEP, 0, D><<<blocks, TB_PLAIN, 0, G->stream>>>(G->geom, G->st, G->d_moves, useMoves, G->d_zob, so, fo); break;
    case 1: games_kernel<DO_STEP, 1, D><<<blocks, TB_PLAIN, 0, G->stream>>>(G->geom, G->st, G->d_moves, useMoves, G->d_zob, so, fo); break;
    case 2: games_kernel<DO_STEP, 2, D><<<blocks, TB_PLAIN, 0, G->stream>>>(G->geom, G->st, G->d_moves, useMoves, G->d_zob, so, fo); break;
    case 3: games_kernel<DO_STEP, 3, D><<<blocks, TB_TILES, 0, G->stream>>>(G->geom, G->st, G->d_moves, useMoves, G->d_zob, so, fo); break;
    default: games_kernel<DO_STEP, 4, D><<<blocks, TB_TILES, 0, G->stream>>>(G->geom, G->st, G->d_moves, useMoves, G->d_zob, so, fo); break;
  }
}
template <bool DO_STEP>
void launchGames(kc_games* G, int feat, int useMoves, const StepOut& so, FeatOut fo) {
  if(G->big) {   // boards beyond 7x7: the general 128-bit kernel (no planes / fp32 NCHW / fp32 NHWC; callers never ask it for trunk tiles)
    const int blocks = (G->geom.numGames + TB_BIG - 1) / TB_BIG;
    games_big_kernel<DO_STEP><<<blocks, TB_BIG, 0, G->stream>>>(G->geom, G->st, static_cast<const BigGeom*>(G->d_bigGeom), G->d_moves, useMoves, G->d_zob, so,
                                                               fo.planes, fo.global, fo.symmetry, fo.permuteDirs, feat > 2 ? 0 : feat);
    G->launches++;
    return;
  }
  int gpb = feat == 3 ? G->geom.NB * 8 : feat == 4 ? G->geom.NB * 2 : TB_PLAIN;   // FEAT 3 / 4: whole trunk tiles per CTA
  fo.gamesPerBlock = gpb;
  int blocks = (G->geom.numGames + gpb - 1) / gpb;
  const Geom& g = G->geom;
  if(g.W == 5 && g.H == 5 && g.K == 4) launchGamesD<DO_STEP, StaticDims<5, 5, 4>>(G, feat, useMoves, so, fo, blocks);
  else if(g.W == 6 && g.H == 6 && g.K == 4 && static6()) launchGamesD<DO_STEP, StaticDims<6, 6, 4>>(G, feat, useMoves, so, fo, blocks);
  else launchGamesD<DO_STEP, DynDims>(G, feat, useMoves, so, fo, blocks);
  G->launches++;
}

StepOut stepOutOf(kc_games* G, bool all) {
  StepOut so;
  so.legal = all ? G->d_legal : nullptr; so.status = all ? G->d_status : nullptr;
  so.sitHash = all ? G->d_sitHash : nullptr; so.played = all ? G->d_played : nullptr;
  so.stats = G->d_stats;
  return so;
}
}  // namespace

namespace kc {
// legal masks, status words and sit-hashes of the current positions into the object's device buffers (no step, no planes)
int gamesRefreshOutputs(kc_games* G) {
  StepOut so = stepOutOf(G, true);
  so.played = nullptr; so.stats = nullptr;
  launchGames<false>(G, 0, 0, so, FeatOut{});
  KC_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace kc

extern "C" {

int kc_games_create(kc_ctx* ctx, int numGames, int xSize, int ySize, int winLen, kc_games** out) {
  KC_CHECK(ctx && out, "kc_games_create: null argument");
  KC_CHECK(numGames > 0, "kc_games_create: numGames must be positive");
  KC_CHECK(xSize >= 2 && ySize >= 2 && xSize <= KC_MAX_LEN && ySize <= KC_MAX_LEN, "kc_games_create: board size must be within 2..10 (board.h:120)");
  KC_CHECK(winLen >= 2 && winLen <= KC_MAX_LEN, "kc_games_create: winLen must be within 2..10");
  // up to 7x7 a padded bitboard (bit = y*(W+1) + x) fits 64 bits: the fast kernels; beyond that the general 128-bit kernel
  const bool big = xSize > KC_MAX_DEVICE_LEN || ySize > KC_MAX_DEVICE_LEN;
  KC_CUDA(cudaSetDevice(ctx->device));
  kc_games* G = new kc_games();
  G->ctx = ctx;
  Geom& g = G->geom;
  memset(&g, 0, sizeof(g));
  g.W = xSize; g.H = ySize; g.K = winLen; g.HW = xSize * ySize; g.stride = xSize + 1;
  g.LW = (4 * g.HW + 31) / 32;
  g.numGames = numGames;
  g.rowMask = (1ULL << g.W) - 1;
  G->big = big;
  if(big) {
    BigGeom bg;
    memset(&bg, 0, sizeof bg);
    for(int y = 0; y < g.H; y++)
      for(int x = 0; x < g.W; x++) {
        const B128 bit = B128::bit(y * g.stride + x);
        bg.all |= bit;
        bg.lines[0][x] |= bit; bg.lines[1][y] |= bit; bg.lines[2][x - y + g.H - 1] |= bit; bg.lines[3][x + y] |= bit;
      }
    KC_CUDA(cudaMalloc(&G->d_bigGeom, sizeof bg));
    KC_CUDA(cudaMemcpy(G->d_bigGeom, &bg, sizeof bg, cudaMemcpyHostToDevice));
  }
  for(int y = 0; y < g.H && !big; y++)
    for(int x = 0; x < g.W; x++) {
      uint64_t bit = 1ULL << (y * g.stride + x);
      g.all |= bit;
      g.lines[0][x] |= bit;
      g.lines[1][y] |= bit;
      g.lines[2][x - y + g.H - 1] |= bit;
      g.lines[3][x + y] |= bit;
    }
  const ZobristTables& z = zobrist();
  memcpy(g.playerHash, z.player, sizeof(g.playerHash));
  g.sizeHash[0] = z.sizeX[g.W][0] ^ z.sizeY[g.H][0];
  g.sizeHash[1] = z.sizeX[g.W][1] ^ z.sizeY[g.H][1];
  g.NB = kc::boardsPerTile(g.W, g.H);
  g.tileRowW = g.NB * g.stride;
  std::vector<uint64_t> zob((size_t)g.HW * 4);
  for(int y = 0; y < g.H; y++)
    for(int x = 0; x < g.W; x++) {
      int spot = (x + 1) + (y + 1) * (g.W + 1);   // board.h:74-75
      for(int c = 0; c < 2; c++) {
        zob[((size_t)(y * g.W + x) * 2 + c) * 2 + 0] = z.board[spot][c + 1][0];
        zob[((size_t)(y * g.W + x) * 2 + c) * 2 + 1] = z.board[spot][c + 1][1];
      }
    }
  size_t n = (size_t)numGames;
  KC_CUDA(cudaMalloc(&G->d_zob, zob.size() * 8));
  KC_CUDA(cudaMemcpy(G->d_zob, zob.data(), zob.size() * 8, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMalloc(&G->st.black, n * 8)); KC_CUDA(cudaMalloc(&G->st.white, n * 8));
  KC_CUDA(cudaMalloc(&G->st.hash0, n * 8)); KC_CUDA(cudaMalloc(&G->st.hash1, n * 8));
  KC_CUDA(cudaMalloc(&G->st.gameId, n * 8)); KC_CUDA(cudaMalloc(&G->st.misc, n * 8));
  if(big) { KC_CUDA(cudaMalloc(&G->st.blackHi, n * 8)); KC_CUDA(cudaMalloc(&G->st.whiteHi, n * 8)); }
  KC_CUDA(cudaMalloc(&G->d_moves, n * 2));
  KC_CUDA(cudaMalloc(&G->d_legal, 4 * n * g.LW * 4)); KC_CUDA(cudaMalloc(&G->d_status, 4 * n * 4));   // 4 slots: the per-ply ring of the multi-ply launches, slot 0 = the current position
  KC_CUDA(cudaMalloc(&G->d_sitHash, 4 * n * 16)); KC_CUDA(cudaMalloc(&G->d_played, 4 * n * 2));
  KC_CUDA(cudaMalloc(&G->d_stats, 8 * 8));
  KC_CUDA(cudaMemset(G->d_stats, 0, 64));
  KC_CUDA(cudaMalloc(&G->d_planes, n * 15 * g.HW * 4)); KC_CUDA(cudaMalloc(&G->d_global, n * 4));
  KC_CUDA(cudaMalloc(&G->d_sym, n)); KC_CUDA(cudaMemset(G->d_sym, 0, n));
  KC_CUDA(cudaStreamCreateWithFlags(&G->stream, cudaStreamNonBlocking));
  KC_CUDA(cudaEventCreate(&G->ev0)); KC_CUDA(cudaEventCreate(&G->ev1));
  *out = G;
  return kc_games_reset(G, 0, 0, 0);
}

int kc_games_destroy(kc_games* G) {
  if(!G) return 0;
  cudaSetDevice(G->ctx->device);
  cudaStreamSynchronize(G->stream);
  cudaFree(G->d_zob); cudaFree(G->st.black); cudaFree(G->st.white); cudaFree(G->st.hash0); cudaFree(G->st.hash1);
  cudaFree(G->st.gameId); cudaFree(G->st.misc); cudaFree(G->d_moves); cudaFree(G->d_legal); cudaFree(G->d_status);
  cudaFree(G->d_sitHash); cudaFree(G->d_played); cudaFree(G->d_stats); cudaFree(G->d_planes); cudaFree(G->d_global);
  cudaFree(G->d_sym); cudaFree(G->d_flush); cudaFree(G->st.blackHi); cudaFree(G->st.whiteHi); cudaFree(G->d_bigGeom);
  cudaFree(G->d_ppPolicy); cudaFree(G->d_ppWinLoss); cudaFree(G->d_ppMisc); cudaFree(G->d_ppHash);
  for(float* p : G->d_planesRing) cudaFree(p);
  for(cudaEvent_t e : G->evPool) cudaEventDestroy(e);
  cudaEventDestroy(G->ev0); cudaEventDestroy(G->ev1);
  cudaStreamDestroy(G->stream);
  delete G;
  return 0;
}

int kc_games_reset(kc_games* G, uint64_t seed, uint64_t firstGameId, int autoRefill) {
  KC_CHECK(G, "kc_games_reset: null games");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  Geom& g = G->geom;
  g.seed = seed;
  g.autoRefill = autoRefill ? 1 : 0;
  size_t n = (size_t)g.numGames;
  std::vector<uint64_t> zero(n, 0), h0(n, g.sizeHash[0]), h1(n, g.sizeHash[1]), ids(n), misc(n, G->big ? BIG_MISC_START : (4ULL << 40) | ((uint64_t)(1 << 3) << 56));
  if(G->big) {
    KC_CUDA(cudaMemcpyAsync(G->st.blackHi, zero.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
    KC_CUDA(cudaMemcpyAsync(G->st.whiteHi, zero.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
  }
  for(size_t i = 0; i < n; i++) ids[i] = firstGameId + i;
  KC_CUDA(cudaMemcpyAsync(G->st.black, zero.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
  KC_CUDA(cudaMemcpyAsync(G->st.white, zero.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
  KC_CUDA(cudaMemcpyAsync(G->st.hash0, h0.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
  KC_CUDA(cudaMemcpyAsync(G->st.hash1, h1.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
  KC_CUDA(cudaMemcpyAsync(G->st.gameId, ids.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
  KC_CUDA(cudaMemcpyAsync(G->st.misc, misc.data(), n * 8, cudaMemcpyHostToDevice, G->stream));
  KC_CUDA(cudaMemsetAsync(G->d_stats, 0, 64, G->stream));
  KC_CUDA(cudaStreamSynchronize(G->stream));
  return 0;
}

int kc_games_load(kc_games* G, int g0, int n, const int8_t* stones, const int8_t* nextPla,
                  const int16_t* moves, const int32_t* numTurns) {
  KC_CHECK(G && stones && nextPla, "kc_games_load: null argument");
  const Geom& g = G->geom;
  KC_CHECK(g0 >= 0 && n > 0 && g0 + n <= g.numGames, "kc_games_load: range out of bounds");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const ZobristTables& z = zobrist();
  std::vector<uint64_t> b(n), w(n), h0(n), h1(n), misc(n);
  if(G->big) {   // 128-bit boards, 9-bit history entries (games_big.cuh)
    std::vector<uint64_t> bh(n), wh(n);
    for(int i = 0; i < n; i++) {
      B128 bb{}, ww{};
      uint64_t a0 = g.sizeHash[0], a1 = g.sizeHash[1];
      for(int y = 0; y < g.H; y++)
        for(int x = 0; x < g.W; x++) {
          const int c = stones[(size_t)i * g.HW + y * g.W + x];
          KC_CHECK(c >= 0 && c <= 2, "kc_games_load: stone colour must be 0, 1 or 2");
          if(c == 0) continue;
          if(c == 1) bb |= B128::bit(y * g.stride + x); else ww |= B128::bit(y * g.stride + x);
          const int spot = (x + 1) + (y + 1) * (g.W + 1);
          a0 ^= z.board[spot][c][0]; a1 ^= z.board[spot][c][1];
        }
      KC_CHECK(nextPla[i] == 1 || nextPla[i] == 2, "kc_games_load: nextPla must be 1 or 2");
      uint64_t m = 0;
      int lastDir = 4;
      if(moves)
        for(int k = 0; k < 5; k++) {
          const int pos = moves[((size_t)i * 5 + (4 - k)) * 2 + 0], pla = moves[((size_t)i * 5 + (4 - k)) * 2 + 1];
          if(pos < 0) continue;
          KC_CHECK(pos < 4 * g.HW && (pla == 1 || pla == 2), "kc_games_load: bad history entry");
          m |= (uint64_t)((pos % g.HW) | (pla << 7)) << (9 * k);
          if(k == 0) lastDir = pos / g.HW;
        }
      const int nt = numTurns ? numTurns[i] : 0;
      KC_CHECK(nt >= 0 && nt <= 255, "kc_games_load: numTurns out of range");
      m |= ((uint64_t)lastDir << 45) | ((uint64_t)nt << 48) | ((uint64_t)(nextPla[i] << 3) << 56);
      b[i] = bb.lo; bh[i] = bb.hi; w[i] = ww.lo; wh[i] = ww.hi; h0[i] = a0; h1[i] = a1; misc[i] = m;
    }
    KC_CUDA(cudaMemcpy(G->st.black + g0, b.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    KC_CUDA(cudaMemcpy(G->st.blackHi + g0, bh.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    KC_CUDA(cudaMemcpy(G->st.white + g0, w.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    KC_CUDA(cudaMemcpy(G->st.whiteHi + g0, wh.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    KC_CUDA(cudaMemcpy(G->st.hash0 + g0, h0.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    KC_CUDA(cudaMemcpy(G->st.hash1 + g0, h1.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    KC_CUDA(cudaMemcpy(G->st.misc + g0, misc.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    return 0;
  }
  for(int i = 0; i < n; i++) {
    uint64_t bb = 0, ww = 0, a0 = g.sizeHash[0], a1 = g.sizeHash[1];
    for(int y = 0; y < g.H; y++)
      for(int x = 0; x < g.W; x++) {
        int c = stones[(size_t)i * g.HW + y * g.W + x];
        KC_CHECK(c >= 0 && c <= 2, "kc_games_load: stone colour must be 0, 1 or 2");
        if(c == 0) continue;
        uint64_t bit = 1ULL << (y * g.stride + x);
        if(c == 1) bb |= bit; else ww |= bit;
        int spot = (x + 1) + (y + 1) * (g.W + 1);
        a0 ^= z.board[spot][c][0]; a1 ^= z.board[spot][c][1];
      }
    KC_CHECK(nextPla[i] == 1 || nextPla[i] == 2, "kc_games_load: nextPla must be 1 or 2");
    uint64_t m = 0;
    int lastDir = 4;
    if(moves) {
      // moves are given oldest first; slot 0 of misc is the most recent
      for(int k = 0; k < 5; k++) {
        int pos = moves[((size_t)i * 5 + (4 - k)) * 2 + 0], pla = moves[((size_t)i * 5 + (4 - k)) * 2 + 1];
        if(pos < 0) continue;
        KC_CHECK(pos < 4 * g.HW && (pla == 1 || pla == 2), "kc_games_load: bad history entry");
        m |= (uint64_t)((pos % g.HW) | (pla << 6)) << (8 * k);
        if(k == 0) lastDir = pos / g.HW;
      }
    }
    int nt = numTurns ? numTurns[i] : 0;
    KC_CHECK(nt >= 0 && nt <= 255, "kc_games_load: numTurns out of range");
    m |= ((uint64_t)lastDir << 40) | ((uint64_t)nt << 48) | ((uint64_t)(nextPla[i] << 3) << 56);
    b[i] = bb; w[i] = ww; h0[i] = a0; h1[i] = a1; misc[i] = m;
  }
  KC_CUDA(cudaMemcpy(G->st.black + g0, b.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMemcpy(G->st.white + g0, w.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMemcpy(G->st.hash0 + g0, h0.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMemcpy(G->st.hash1 + g0, h1.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMemcpy(G->st.misc + g0, misc.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  return 0;
}

int kc_games_step(kc_games* G, const int16_t* movePos, uint32_t* legal, uint32_t* status,
                  uint64_t* sitHash, int16_t* played, uint64_t* gameIds) {
  KC_CHECK(G, "kc_games_step: null games");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const Geom& g = G->geom;
  size_t n = (size_t)g.numGames;
  if(movePos) KC_CUDA(cudaMemcpyAsync(G->d_moves, movePos, n * 2, cudaMemcpyHostToDevice, G->stream));
  launchGames<true>(G, 0, movePos ? 1 : 0, stepOutOf(G, true), FeatOut{});
  KC_CUDA(cudaGetLastError());
  if(legal) KC_CUDA(cudaMemcpyAsync(legal, G->d_legal, n * g.LW * 4, cudaMemcpyDeviceToHost, G->stream));
  if(status) KC_CUDA(cudaMemcpyAsync(status, G->d_status, n * 4, cudaMemcpyDeviceToHost, G->stream));
  if(sitHash) KC_CUDA(cudaMemcpyAsync(sitHash, G->d_sitHash, n * 16, cudaMemcpyDeviceToHost, G->stream));
  if(played) KC_CUDA(cudaMemcpyAsync(played, G->d_played, n * 2, cudaMemcpyDeviceToHost, G->stream));
  if(gameIds) KC_CUDA(cudaMemcpyAsync(gameIds, G->st.gameId, n * 8, cudaMemcpyDeviceToHost, G->stream));
  KC_CUDA(cudaStreamSynchronize(G->stream));
  return 0;
}

int kc_games_features(kc_games* G, int layout, const int8_t* symmetry, float* planes, float* global) {
  KC_CHECK(G && planes, "kc_games_features: null argument");
  KC_CHECK(layout == 0 || layout == 1, "kc_games_features: layout must be 0 (NCHW) or 1 (NHWC)");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const Geom& g = G->geom;
  size_t n = (size_t)g.numGames;
  if(symmetry) {
    for(size_t i = 0; i < n; i++) KC_CHECK(symmetry[i] >= 0 && symmetry[i] < 8, "kc_games_features: symmetry out of range");
    KC_CUDA(cudaMemcpyAsync(G->d_sym, symmetry, n, cudaMemcpyHostToDevice, G->stream));
  }
  FeatOut fo{};
  fo.planes = G->d_planes; fo.global = G->d_global; fo.symmetry = symmetry ? G->d_sym : nullptr;
  StepOut so{};
  launchGames<false>(G, layout == 0 ? 1 : 2, 0, so, fo);
  KC_CUDA(cudaGetLastError());
  KC_CUDA(cudaMemcpyAsync(planes, G->d_planes, n * 15 * g.HW * 4, cudaMemcpyDeviceToHost, G->stream));
  if(global) KC_CUDA(cudaMemcpyAsync(global, G->d_global, n * 4, cudaMemcpyDeviceToHost, G->stream));
  KC_CUDA(cudaStreamSynchronize(G->stream));
  return 0;
}

int kc_games_eval(kc_games* G, kc_handle* h, const int8_t* symmetry) { return kc::gamesEval(G, h, symmetry, nullptr); }

}  // extern "C"

namespace kc {
int gamesEval(kc_games* G, kc_handle* h, const int8_t* symmetry, const int* nDev, int rowOffset, bool smallCtas, bool symOnDevice) {
  KC_CHECK(G && h, "kc_games_eval: null argument");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const Geom& g = G->geom;
  size_t n = (size_t)g.numGames;
  if(kc::handleCheckGeometry(h, g.W, g.H, g.numGames + rowOffset)) return 1;
  KC_CHECK(rowOffset == 0 || (kc::handleIsBf16(h) && rowOffset % (2 * g.NB) == 0), "kc_games_eval: a row offset needs the bf16 path and whole work items");
  if(symmetry) KC_CUDA(cudaMemcpyAsync(G->d_sym, symmetry, n, cudaMemcpyHostToDevice, G->stream));
  const bool haveSym = symmetry != nullptr || symOnDevice;   // symOnDevice: d_sym was filled by a kernel (the search's nnRandomize)
  FeatOut fo{};
  fo.symmetry = haveSym ? G->d_sym : nullptr;
  fo.permuteDirs = (haveSym && kc::handlePermutesDirs(h)) ? 1 : 0;
  StepOut so = stepOutOf(G, true);   // also refreshes legal masks / status / sit-hashes of the evaluated positions
  so.played = nullptr; so.stats = nullptr;
  if(kc::handleIsBf16(h) && G->big) {
    // boards beyond 7x7: plain fp32 NCHW planes into the handle's raw input rows, then the conversion kc_forward uses (it applies
    // the symmetry and builds the tiles)
    KC_CHECK(rowOffset == 0 && !nDev, "kc_games_eval: boards beyond 7x7 are evaluated as one plain batch");
    fo.planes = kc::handleRawInput(h); fo.global = kc::handleRawGlobal(h); fo.symmetry = nullptr; fo.permuteDirs = 0;
    launchGames<false>(G, 1, 0, so, fo);
    KC_CUDA(cudaGetLastError());
    if(kc::handleConvertRaw(h, g.numGames, haveSym ? G->d_sym : nullptr, G->stream)) return 1;
  } else if(kc::handleIsBf16(h)) {
    fo.tiles = (uint4*)kc::handleInputTiles(h) + (size_t)(rowOffset / g.NB) * 2 * TILE_ROWS;
    kc::handleTileConstants(h, (float)g.K, &fo.tileOne, &fo.tileK);
    launchGames<false>(G, smallCtas ? 4 : 3, 0, so, fo);
  } else {
    fo.planes = kc::handleInputNHWC(h); fo.global = kc::handleInputGlobal(h);
    launchGames<false>(G, 2, 0, so, fo);
  }
  KC_CUDA(cudaGetLastError());
  return kc::handleRunOnStream(h, g.numGames, G->stream, haveSym ? G->d_sym : nullptr, nDev, rowOffset, /*symIsLocal=*/true);
}
}  // namespace kc

extern "C" {

int kc_games_run_timed(kc_games* G, kc_handle* h, int plies, size_t flushL2Bytes, kc_stats* acc, float* msTotal) {
  KC_CHECK(G && plies > 0, "kc_games_run: bad argument");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const Geom& g = G->geom;
  if(h && kc::handleCheckGeometry(h, g.W, g.H, g.numGames)) return 1;
  if(flushL2Bytes > G->flushBytes) {
    cudaFree(G->d_flush);
    G->d_flush = nullptr; G->flushBytes = 0;
    KC_CUDA(cudaMalloc(&G->d_flush, flushL2Bytes));
    G->flushBytes = flushL2Bytes;
  }
  while((int)G->evPool.size() < 2 * plies) {
    cudaEvent_t e;
    KC_CUDA(cudaEventCreate(&e));
    G->evPool.push_back(e);
  }
  KC_CUDA(cudaMemsetAsync(G->d_stats, 0, 64, G->stream));
  // Rules+features only: plies are timed in groups of RING launches between one event pair (a ~25 us kernel is too short
  // for a per-launch event pair: the pair itself costs a few us), each launch of a group writing its planes to a
  // different ring slot.  With a net: one event pair per ply.
  constexpr int RING = 4;
  const int grp = h ? 1 : RING;
  if(!h && flushL2Bytes)
    for(int i = 0; i < RING - 1; i++)
      if(!G->d_planesRing[i]) KC_CUDA(cudaMalloc(&G->d_planesRing[i], (size_t)g.numGames * 15 * g.HW * 4));
  int nGroups = 0;
  if(!h && G->big) {
    // boards beyond 7x7: the general kernel, one launch (rules + fp32 NCHW planes) and one event pair per ply
    for(int p = 0; p < plies; p++) {
      if(flushL2Bytes) KC_CUDA(cudaMemsetAsync(G->d_flush, p & 0xff, flushL2Bytes, G->stream));
      KC_CUDA(cudaEventRecord(G->evPool[2 * nGroups], G->stream));
      FeatOut fo{};
      fo.planes = G->d_planes; fo.global = G->d_global;
      launchGames<true>(G, 1, 0, stepOutOf(G, true), fo);
      KC_CUDA(cudaEventRecord(G->evPool[2 * nGroups + 1], G->stream));
      nGroups++;
    }
    G->lastRunPlanes = G->d_planes; G->lastRunPlies = 1; G->lastRunRing = false;
  } else if(!h) {
    // rules + features only: one multi-ply launch per group of RING plies, one event pair per launch
    PlaneRing ring;
    for(int i = 0; i < RING; i++) ring.slot[i] = (flushL2Bytes && i > 0) ? G->d_planesRing[i - 1] : G->d_planes;
    // every ply's masks / status words / hashes / moves get their own slot too: the object's buffers hold 4 slots of G entries, slot 0 first
    const StepOut so = stepOutOf(G, true);
    const int blocks = (g.numGames + TB_PLAIN - 1) / TB_PLAIN;
    constexpr int PLIES_PER_LAUNCH = 8;   // two turns of the 4-slot ring (412 MB > L2 between two writes of a slot)
    for(int p = 0; p < plies; p += PLIES_PER_LAUNCH) {
      const int np = std::min(PLIES_PER_LAUNCH, plies - p);
      if(flushL2Bytes) KC_CUDA(cudaMemsetAsync(G->d_flush, p & 0xff, flushL2Bytes, G->stream));
      KC_CUDA(cudaEventRecord(G->evPool[2 * nGroups], G->stream));
      const int first = (RING - ((np - 1) & 3)) & 3;   // the launch's last ply writes slot 0
      // producer / consumer warps for the static 5x5 instantiation (5.72 -> 6.10 TB/s algorithmic at 65,536 games); the generic
      // 64-bit kernel is bound by its rolled rules code and got slower when split (6x6: 4.46 -> 3.83 TB/s), so it keeps one warp
      // per 32 games.  KC_GAMES_MULTI_SPLIT = 0 / 1 forces one form for every board.
      static const int splitEnv = [] { const char* e = getenv("KC_GAMES_MULTI_SPLIT"); return e ? atoi(e) : -1; }();
      const bool static5 = g.W == 5 && g.H == 5 && g.K == 4;
      const bool stat6 = g.W == 6 && g.H == 6 && g.K == 4 && static6();
      if(stat6) {
        if(splitEnv != 0)
          games_multi_split_kernel<StaticDims<6, 6, 4>, 1><<<blocks, 2 * TB_PLAIN, 0, G->stream>>>(g, G->st, G->d_zob, so, ring, G->d_global, np, first);
        else
          games_multi_kernel<StaticDims<6, 6, 4>><<<blocks, TB_PLAIN, 0, G->stream>>>(g, G->st, G->d_zob, so, ring, G->d_global, np, first);
      } else if(splitEnv < 0 ? static5 : splitEnv != 0) {
        if(static5 && splitEnv == 2)
          games_multi_split_kernel<StaticDims<5, 5, 4>, 2><<<blocks, 3 * TB_PLAIN, 0, G->stream>>>(g, G->st, G->d_zob, so, ring, G->d_global, np, first);
        else if(static5)
          games_multi_split_kernel<StaticDims<5, 5, 4>, 1><<<blocks, 2 * TB_PLAIN, 0, G->stream>>>(g, G->st, G->d_zob, so, ring, G->d_global, np, first);
        else
          games_multi_split_kernel<DynDims, 1><<<blocks, 2 * TB_PLAIN, 0, G->stream>>>(g, G->st, G->d_zob, so, ring, G->d_global, np, first);
      } else if(g.W == 5 && g.H == 5 && g.K == 4)
        games_multi_kernel<StaticDims<5, 5, 4>><<<blocks, TB_PLAIN, 0, G->stream>>>(g, G->st, G->d_zob, so, ring, G->d_global, np, first);
      else
        games_multi_kernel<DynDims><<<blocks, TB_PLAIN, 0, G->stream>>>(g, G->st, G->d_zob, so, ring, G->d_global, np, first);
      G->launches++;
      KC_CUDA(cudaEventRecord(G->evPool[2 * nGroups + 1], G->stream));
      nGroups++;
      G->lastRunPlanes = ring.slot[0];
      G->lastRunPlies = np; G->lastRunRing = flushL2Bytes != 0;
    }
  }
  for(int p = 0; h && p < plies; p++) {
    // optional L2 flush between timed windows, outside them
    if(p % grp == 0) {
      if(flushL2Bytes) KC_CUDA(cudaMemsetAsync(G->d_flush, p & 0xff, flushL2Bytes, G->stream));
      KC_CUDA(cudaEventRecord(G->evPool[2 * nGroups], G->stream));
    }
    FeatOut fo{};
    StepOut so = stepOutOf(G, true);
    if(!h) {
      const int slot = flushL2Bytes ? p % RING : 0;
      fo.planes = slot == 0 ? G->d_planes : G->d_planesRing[slot - 1]; fo.global = G->d_global;
      launchGames<true>(G, 1, 0, so, fo);
    } else if(kc::handleIsBf16(h) && G->big) {   // boards beyond 7x7: raw fp32 rows + the conversion kc_forward uses (see gamesEval)
      fo.planes = kc::handleRawInput(h); fo.global = kc::handleRawGlobal(h);
      launchGames<true>(G, 1, 0, so, fo);
      if(kc::handleConvertRaw(h, g.numGames, nullptr, G->stream)) return 1;
      if(kc::handleRunOnStream(h, g.numGames, G->stream, nullptr)) return 1;
    } else if(kc::handleIsBf16(h)) {
      fo.tiles = (uint4*)kc::handleInputTiles(h);
      kc::handleTileConstants(h, (float)g.K, &fo.tileOne, &fo.tileK);
      launchGames<true>(G, 3, 0, so, fo);
      if(kc::handleRunOnStream(h, g.numGames, G->stream, nullptr)) return 1;
    } else {
      fo.planes = kc::handleInputNHWC(h); fo.global = kc::handleInputGlobal(h);
      launchGames<true>(G, 2, 0, so, fo);
      if(kc::handleRunOnStream(h, g.numGames, G->stream, nullptr)) return 1;
    }
    if(p % grp == grp - 1 || p == plies - 1) { KC_CUDA(cudaEventRecord(G->evPool[2 * nGroups + 1], G->stream)); nGroups++; }
  }
  KC_CUDA(cudaGetLastError());
  unsigned long long hs[8];
  KC_CUDA(cudaMemcpyAsync(hs, G->d_stats, 64, cudaMemcpyDeviceToHost, G->stream));
  KC_CUDA(cudaStreamSynchronize(G->stream));
  float msSum = 0.f;
  for(int p = 0; p < nGroups; p++) {
    float ms = 0.f;
    KC_CUDA(cudaEventElapsedTime(&ms, G->evPool[2 * p], G->evPool[2 * p + 1]));
    msSum += ms;
  }
  if(!h) G->lastKernelMs = msSum / plies;
  if(msTotal) *msTotal = msSum;
  if(h && kc::handleCheckAbort(h)) return 1;
  if(acc) {
    acc->steps += hs[0];
    acc->evals += h ? (uint64_t)g.numGames * plies : 0;   // every lane's position goes through the net each ply
    acc->gamesFinished += hs[2]; acc->blackWins += hs[3]; acc->whiteWins += hs[4]; acc->draws += hs[5];
    acc->checksum ^= hs[6];
  }
  return 0;
}

int kc_games_read_run_ply(kc_games* G, int pliesBack, float* planes, uint32_t* legal, uint32_t* status, uint64_t* sitHash, int16_t* played) {
  KC_CHECK(G, "kc_games_read_run_ply: null argument");
  KC_CHECK(G->lastRunPlanes, "kc_games_read_run_ply: call kc_games_run without a handle first");
  KC_CHECK(pliesBack >= 0 && pliesBack < 4 && pliesBack < G->lastRunPlies, "kc_games_read_run_ply: pliesBack must be below min(4, plies of the last launch)");
  KC_CHECK(!planes || pliesBack == 0 || G->lastRunRing, "kc_games_read_run_ply: earlier plies' planes are kept only when the run used the plane ring (flushL2Bytes > 0)");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const Geom& g = G->geom;
  const size_t n = (size_t)g.numGames;
  const int slot = (4 - pliesBack) & 3;
  KC_CUDA(cudaStreamSynchronize(G->stream));
  if(planes) KC_CUDA(cudaMemcpy(planes, slot ? G->d_planesRing[slot - 1] : G->d_planes, n * 15 * g.HW * 4, cudaMemcpyDeviceToHost));
  if(sitHash) KC_CUDA(cudaMemcpy(sitHash, G->d_sitHash + 2 * n * slot, n * 16, cudaMemcpyDeviceToHost));
  if(legal) KC_CUDA(cudaMemcpy(legal, G->d_legal + n * g.LW * slot, n * g.LW * 4, cudaMemcpyDeviceToHost));
  if(status) KC_CUDA(cudaMemcpy(status, G->d_status + n * slot, n * 4, cudaMemcpyDeviceToHost));
  if(played) KC_CUDA(cudaMemcpy(played, G->d_played + n * slot, n * 2, cudaMemcpyDeviceToHost));
  return 0;
}

int kc_games_read_run_outputs(kc_games* G, float* planes, float* global, uint32_t* legal, uint32_t* status, uint64_t* sitHash, int16_t* played) {
  KC_CHECK(G, "kc_games_read_run_outputs: null argument");
  KC_CHECK(G->lastRunPlanes, "kc_games_read_run_outputs: call kc_games_run without a handle first");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const Geom& g = G->geom;
  const size_t n = (size_t)g.numGames;
  KC_CUDA(cudaStreamSynchronize(G->stream));
  if(planes) KC_CUDA(cudaMemcpy(planes, G->lastRunPlanes, n * 15 * g.HW * 4, cudaMemcpyDeviceToHost));
  if(global) KC_CUDA(cudaMemcpy(global, G->d_global, n * 4, cudaMemcpyDeviceToHost));
  if(legal) KC_CUDA(cudaMemcpy(legal, G->d_legal, n * g.LW * 4, cudaMemcpyDeviceToHost));
  if(status) KC_CUDA(cudaMemcpy(status, G->d_status, n * 4, cudaMemcpyDeviceToHost));
  if(sitHash) KC_CUDA(cudaMemcpy(sitHash, G->d_sitHash, n * 16, cudaMemcpyDeviceToHost));
  if(played) KC_CUDA(cudaMemcpy(played, G->d_played, n * 2, cudaMemcpyDeviceToHost));
  return 0;
}

int kc_games_run(kc_games* G, kc_handle* h, int plies, kc_stats* acc) {
  return kc_games_run_timed(G, h, plies, 0, acc, nullptr);
}

int kc_games_postprocess(kc_games* G, kc_handle* h, float policyTemperature, float* policyProbs, float* whiteWinLoss, float* misc,
                         uint64_t* nnHash) {
  KC_CHECK(G && h && policyTemperature > 0.f, "kc_games_postprocess: bad argument");
  KC_CUDA(cudaSetDevice(G->ctx->device));
  const Geom& g = G->geom;
  if(kc::handleCheckGeometry(h, g.W, g.H, g.numGames)) return 1;
  size_t n = (size_t)g.numGames;
  if(!G->d_ppPolicy) {   // owned by the games object: no allocation (and no implicit device synchronisation) per call
    KC_CUDA(cudaMalloc(&G->d_ppPolicy, n * 4 * g.HW * 4)); KC_CUDA(cudaMalloc(&G->d_ppWinLoss, n * 8));
    KC_CUDA(cudaMalloc(&G->d_ppMisc, n * 8)); KC_CUDA(cudaMalloc(&G->d_ppHash, n * 16));
  }
  float *dP = G->d_ppPolicy, *dV = G->d_ppWinLoss, *dM = G->d_ppMisc; uint64_t* dH = G->d_ppHash;
  kc::launchPostprocess(h, g.numGames, g.LW, G->d_legal, G->d_status, G->d_sitHash, policyTemperature, dP, dV, dM, dH, G->stream);
  G->launches++;
  KC_CUDA(cudaGetLastError());
  if(policyProbs) KC_CUDA(cudaMemcpyAsync(policyProbs, dP, n * 4 * g.HW * 4, cudaMemcpyDeviceToHost, G->stream));
  if(whiteWinLoss) KC_CUDA(cudaMemcpyAsync(whiteWinLoss, dV, n * 8, cudaMemcpyDeviceToHost, G->stream));
  if(misc) KC_CUDA(cudaMemcpyAsync(misc, dM, n * 8, cudaMemcpyDeviceToHost, G->stream));
  if(nnHash) KC_CUDA(cudaMemcpyAsync(nnHash, dH, n * 16, cudaMemcpyDeviceToHost, G->stream));
  KC_CUDA(cudaStreamSynchronize(G->stream));
  return 0;
}

int64_t kc_games_launch_count(const kc_games* G) { return G ? G->launches : 0; }
float kc_games_last_kernel_ms(const kc_games* G) { return G ? G->lastKernelMs : 0.f; }

}  // extern "C"
