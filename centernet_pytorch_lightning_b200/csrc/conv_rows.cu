// Row-window convolution for the wide, thin layers (W_out % 128 == 0, Ci in {8,16,32,64}) on tcgen05.
//
// Replaces (reference file:line under CenterNet/models/backbones/pose_dla_dcn.py): the stem :281-285 (7x7, 3->16
// @512^2), the conv levels :351-370 (16->16 @512^2, 16->32 stride 2), the stride-2 / stride-1 3x3 convs of the first
// Tree (BasicBlock :28-68, 32->64 and 64->64 @128^2) and the 64->27 conv_offset_mask convs of the DCN layers that
// run at 128^2 (pose_dla_dcn.py:441) -- each with BatchNorm(eval) / bias, residual and ReLU in the epilogue.
//
// Why another kernel: with few channels the implicit-GEMM K dimension is short (144..576) and N is tiny
// (16..64), so a tile's MMA time is a few hundred clocks while an im2col feed re-reads every input pixel KH*KW
// times from L2.  Here every input pixel is fetched ONCE per 128-pixel column strip and the im2col structure is
// expressed purely through shared-memory *descriptors*:
//   * a CTA walks down a strip (image n, 128 output columns) row by row and keeps the last KH input rows in a
//     ring of shared-memory slots; one slot = one input row, stored as planes [phase][8-channel chunk][pixel] of
//     16-byte pixels (phase = x mod stride, so a stride-2 conv also reads unit-stride planes);
//   * in the canonical no-swizzle K-major layout a core matrix is 8 rows x 16 B contiguous, i.e. 8 consecutive
//     pixels of one plane.  The A operand of filter tap (kh, kw) is therefore just the plane of row slot kh,
//     started kw pixels later: descriptor start = plane + kw*16 B, SBO = 128 B (next 8 pixels), LBO = the distance
//     to the second 8-channel half of the K=16 step (the next chunk plane, or -- 8-channel stem -- the next pixel,
//     which pairs the taps kw, kw+1).  No im2col copy exists anywhere, not even in shared memory;
//   * the packed weights (all of K x Co) stay resident in shared memory for the CTA's lifetime (TMA, SWIZZLE_128B).
// Warp roles: 4 producer warps (16-byte cp.async with zero fill at the borders, completion handed to an mbarrier
// two rows behind), 1 MMA warp (one lane issues tcgen05.mma M=128, N=Co, K=16), 4 epilogue warps; two TMEM
// accumulators so the epilogue of row i overlaps the MMAs of row i+1.
#include "umma.cuh"
#include "tma_host.h"
#include <stdlib.h>

namespace cnb {
namespace {

// CNB_ROWS_DEBUG stage-skipping experiments: compiled in only with -DCNB_ROWS_EXPERIMENTS (their predicates sit in
// the role loops)
#ifdef CNB_ROWS_EXPERIMENTS
#define ROWS_DBG(a) ((a).debug)
#else
#define ROWS_DBG(a) 0
#endif

constexpr int BM = 128;
constexpr int NPW = 8;                       // producer warps
constexpr int NPT = NPW * 32;
constexpr int MAXI = 5;                      // 16-byte copies per producer thread and input row (<= 1280 per row)
constexpr int NMW = 8;                       // MMA warps: one per output row of a unit (independent issue streams)
constexpr int W_MMA = NPW;                   // warps 8..15
constexpr int W_EPI0 = NPW + NMW;            // warps 16..23: TMEM lane quarter = warp & 3, two warps per quarter
constexpr int NEW = 8;                       // epilogue warps (the (row, 16-channel group) items of a unit alternate
                                             // between the two warps of a quarter)
constexpr int NTHREADS = (W_EPI0 + NEW) * 32;  // 768
constexpr int MAX_SLOTS = 48;
constexpr int MAX_ACC = 8;                   // TMEM accumulators in rotation (epilogue latency hiding)
constexpr int MAX_STEPS = 64;                // K=16 steps per output tile
constexpr int MAX_DEPTH = 24;                // input rows a producer thread may keep in flight (cp.async groups)

struct RArgs {
  cnb_conv_desc d;
  const __nv_bfloat16* x;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;
  void* y;
  int s;             // stride (1 or 2)
  int nch;           // Ci / 8: chunk planes per phase
  int KWp;           // KW as packed in the weights (KW rounded up to even when Ci == 8)
  int PW;            // pixels stored per plane (128 + halo)
  int q_off;         // plane pixel j holds q = ox0 + q_off + j   (input x = s*q + phase)
  u32 plane_bytes;
  u32 slot_bytes;
  int nslots;
  int depth;         // input rows in flight per producer thread (<= MAX_DEPTH)
  int npw;           // producer warps that take part (rows with few 16-byte cells do not need all NPW)
  int nk;            // cells per participating producer thread and row (<= MAXI)
  int nseg;          // Wo / 128
  int R;             // output rows per unit (accumulators interleaved by the MMA warp)
  int upr;           // units per strip = ceil(Ho / R)
  int units;         // B * nseg * upr
  int BN;
  int nslab;         // 64-wide K slabs of the resident weights
  u32 b_slab_bytes, b_bytes;
  u32 tmem_cols, acc_stride, idesc;
  int nbuf;          // unit buffers in TMEM (each R accumulators), power of two
  int nbuf_sh;
  int spk;           // K=16 steps per filter row
  u32 aoff16[16];
  int debug;               // CNB_ROWS_DEBUG bit mask (timing experiments only): 1 no input copies, 2 no stores, 4 no MMAs    // per step of a filter row: offset of its A window inside a row slot, in 16-byte units
};

__device__ __forceinline__ void cp_async16(u32 dst, const void* src, u32 src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int n) {   // at most n groups still pending
#define CNB_CASE(N) case N: cp_async_wait<N>(); break;
  switch (n) {
    CNB_CASE(0) CNB_CASE(1) CNB_CASE(2) CNB_CASE(3) CNB_CASE(4) CNB_CASE(5) CNB_CASE(6) CNB_CASE(7)
    CNB_CASE(8) CNB_CASE(9) CNB_CASE(10) CNB_CASE(11) CNB_CASE(12) CNB_CASE(13) CNB_CASE(14) CNB_CASE(15)
    CNB_CASE(16) CNB_CASE(17) CNB_CASE(18) CNB_CASE(19) CNB_CASE(20) CNB_CASE(21) CNB_CASE(22)
    default: cp_async_wait<23>(); break;
  }
#undef CNB_CASE
}

// Walks the CTA's units in order.  A unit = up to R consecutive output rows of one strip (their K-step chains
// are interleaved on R accumulators: back-to-back MMAs into one accumulator are latency bound, ~120 clocks
// each, however small N is).  The unit's input rows i = 0..cnt-1 have ring slots sb+i; the walker tells which of
// them are new.  Used identically by the producer and the MMA warp so that both agree on every row's slot.
struct RowWalk {
  int Ho, s, KH, R, upr, nslots;
  int strip, ug;       // current unit: strip and unit index inside the strip
  int nr, cnt;         // output rows of the unit, input rows it needs ((nr-1)*s + KH)
  int sb;              // ring slot of the unit's first input row (iy0 = ug*R*s - pad)
  u32 pb;              // parity of how often the ring has wrapped at that row
  bool fresh;          // first unit of a run (strip start): all rows are new
  __device__ void set_rows() {
    nr = min(R, Ho - ug * R);
    cnt = (nr - 1) * s + KH;
  }
  __device__ void start(int u, int Ho_, int s_, int KH_, int R_, int upr_, int nslots_) {
    Ho = Ho_; s = s_; KH = KH_; R = R_; upr = upr_; nslots = nslots_;
    strip = u / upr;
    ug = u - strip * upr;
    sb = 0;
    pb = 0;
    fresh = true;
    set_rows();
  }
  __device__ int first_new() const { return fresh ? 0 : cnt - nr * s; }   // rows [first_new, cnt) are new
  __device__ void slot_of(int i, int* slot, u32* par) const {             // i < nslots
    int sl = sb + i;
    u32 p = pb;
    if (sl >= nslots) {
      sl -= nslots;
      p ^= 1u;
    }
    *slot = sl;
    *par = p;
  }
  __device__ bool last_of_run(int u, int u_end) const { return u + 1 == u_end || ug == upr - 1; }
  // input rows that no later unit needs once this one is done
  __device__ int released(int u, int u_end) const { return last_of_run(u, u_end) ? cnt : R * s; }
  __device__ void next() {   // advance to the following unit
    int adv;
    if (ug == upr - 1) {     // next strip: every row is reloaded
      adv = cnt;
      ++strip;
      ug = 0;
      fresh = true;
    } else {
      adv = R * s;
      ++ug;
      fresh = false;
    }
    sb += adv;
    while (sb >= nslots) {
      sb -= nslots;
      pb ^= 1u;
    }
    set_rows();
  }
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tmB, const RArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_full[MAX_SLOTS];
  __shared__ __align__(8) u64 s_empty[MAX_SLOTS];
  __shared__ __align__(8) u64 s_tfull[MAX_ACC];
  __shared__ __align__(8) u64 s_tempty[MAX_ACC];
  __shared__ __align__(8) u64 s_bfull;
  __shared__ u32 s_tmem;

  const cnb_conv_desc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const u32 ring_base = smem_base + a.b_bytes;
  float* s_scale = reinterpret_cast<float*>(smem_dyn + (smem_base - smem_u32(smem_dyn)) + a.b_bytes +
                                            (size_t)a.nslots * a.slot_bytes);
  float* s_shift = s_scale + a.BN;

  if (tid == 0) {
    for (int i = 0; i < a.nslots; ++i) {
      mbar_init(&s_full[i], 1);         // the one producer warp that copied the row
      mbar_init(&s_empty[i], 1);        // released by epilogue warp 0 once the unit's accumulators are complete
    }
    for (int i = 0; i < MAX_ACC; ++i) {
      mbar_init(&s_tfull[i], NMW);
      mbar_init(&s_tempty[i], NEW);
    }
    mbar_init(&s_bfull, 1);
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc(&s_tmem, a.tmem_cols);
  pdl_launch_dependents();
  pdl_wait();   // global memory (input rows, residual, scale/shift) is read only after the previous kernel has completed
  for (int i = tid; i < a.BN; i += NTHREADS) {
    s_scale[i] = (i < d.Co && a.scale) ? a.scale[i] : 1.f;
    s_shift[i] = (i < d.Co && a.shift) ? a.shift[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = s_tmem;

  const int u_begin = (int)((long long)blockIdx.x * a.units / gridDim.x);
  const int u_end = (int)((long long)(blockIdx.x + 1) * a.units / gridDim.x);
  if (u_begin >= u_end) {
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem_base, a.tmem_cols);
    return;
  }

  if (warp < NPW) {
    if (warp >= a.npw) goto done;   // small rows: fewer producer warps, fewer arrivals per row
    // =============================== producers: input rows -> ring slots =====================================
    RowWalk w;
    w.start(u_begin, d.Ho, a.s, d.KH, a.R, a.upr, a.nslots);
    const int items = a.s * a.nch * a.PW;
    const int nch_sh = a.nch == 1 ? 0 : (a.nch == 2 ? 1 : (a.nch == 4 ? 2 : 3));
    const int per_phase = a.nch * a.PW;
    // Input rows are dealt out to the producer warps round-robin: ONE warp copies a whole row (lane-strided
    // 16-byte cp.async, zero fill outside the image) and alone arrives on the slot's full barrier.  The per-row
    // bookkeeping (slot wait, commit, wait_group, proxy fence, arrive) is a serial chain of ~100 instructions per
    // warp; with every warp handling every row it capped the whole kernel at one row per ~500 clocks.  A row's
    // arrival happens `wdepth - 1` of the warp's own rows later, so that many rows stay in flight per warp.
    // the rows a warp still has to ISSUE before it arrives for an earlier one must fit in the ring's slack:
    // npw * (wdepth - 1) <= depth, or the MMA warps would wait for an arrival that waits for a free slot
    const int wdepth = min(8, 1 + a.depth / a.npw);
    int pend[8];           // slots of this warp's rows whose copies are in flight (oldest first)
    int npend = 0;
    auto complete_oldest = [&]() {   // the oldest row's copies have landed (caller waited on the group)
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[pend[0]]);
#pragma unroll
      for (int i = 1; i < 8; ++i) pend[i - 1] = pend[i];
      --npend;
    };
    int n = w.strip / a.nseg, seg = w.strip - n * a.nseg;
    int turn = 0;          // running row counter mod npw (identical in every producer warp): whose row this is
    for (int u = u_begin; u < u_end; ++u) {
      const int q0 = seg * BM + a.q_off;              // q of plane pixel 0
      const int iy0 = w.ug * a.R * a.s - d.pad;
      for (int i = w.first_new(); i < w.cnt; ++i) {
        const bool mine = turn == warp;
        if (++turn == a.npw) turn = 0;
        if (!mine) continue;
        const int iy = iy0 + i;
        int slot;
        u32 par;
        w.slot_of(i, &slot, &par);
        mbar_wait_parked(&s_empty[slot], par ^ 1u);
        const u32 dst0 = ring_base + (u32)slot * a.slot_bytes;
        const bool row_ok = iy >= 0 && iy < d.Hi;
        const __nv_bfloat16* rowp = a.x + ((size_t)(n * d.Hi + (row_ok ? iy : 0)) * d.Wi) * d.x_cstride + d.x_coffset;
        for (int it = lane; it < ((ROWS_DBG(a) & 1) ? 0 : items); it += 32) {
          const int p = it >= per_phase ? 1 : 0;       // stride <= 2: at most two phases
          const int r = it - p * per_phase;
          const int c = r & (a.nch - 1);
          const int j = r >> nch_sh;
          const int xg = a.s * (q0 + j) + p;
          const bool ok = row_ok && xg >= 0 && xg < d.Wi;
          cp_async16(dst0 + (u32)(p * a.nch + c) * a.plane_bytes + (u32)j * 16u,
                     ok ? rowp + (size_t)xg * d.x_cstride + c * 8 : a.x, ok ? 16u : 0u);
        }
        cp_async_commit();
        pend[npend++] = slot;
        if (npend == wdepth) {
          cp_async_wait_dyn(wdepth - 1);
          complete_oldest();
        }
      }
      if (w.ug == a.upr - 1 && ++seg == a.nseg) {   // the next unit starts another strip
        seg = 0;
        ++n;
      }
      w.next();
    }
    while (npend > 0) {   // drain
      cp_async_wait_dyn(npend - 1);
      complete_oldest();
    }
  } else if (warp < W_EPI0) {
    // =============================== MMA issuers (warp W_MMA + j drives output row j of every unit) =========== ==============================================================
    // Descriptors are pure integer arithmetic on kernel parameters and loop counters (uniform datapath): the
    // per-step A offsets come from the parameter block (host-computed), nothing is loaded from memory between
    // two tcgen05.mma.
    const int ncp = a.nch >= 2 ? a.nch / 2 : 0;
    const int steps_per_kh = a.spk;
    {   // all lanes walk the loop (warp-uniform operands -> uniform registers); one elected lane issues
      const int j = warp - W_MMA;
      if (j == 0 && elect_one()) {   // resident weights
        mbar_expect_tx(&s_bfull, a.b_bytes);
        for (int sl = 0; sl < a.nslab; ++sl)
          tma_load_2d(smem_base + (u32)sl * a.b_slab_bytes, &tmB, sl * 64, 0, &s_bfull);
      }
      __syncwarp();
      mbar_wait_parked(&s_bfull, 0);
      RowWalk w;
      w.start(u_begin, d.Ho, a.s, d.KH, a.R, a.upr, a.nslots);
      const u32 lbo = ncp ? a.plane_bytes : 16u;
      const u64 dslot0 = make_sdesc(ring_base, lbo, 128, 0);   // slot i adds i * slot_bytes / 16
      const u32 slot16 = a.slot_bytes >> 4;
      const u64 db0 = make_sdesc(smem_base, 16, 1024, 2);      // K slab j adds j * b_slab_bytes / 16, step +2
      const u32 bslab16 = a.b_slab_bytes >> 4;
      u32 t = 0;
      for (int u = u_begin; u < u_end; ++u, ++t) {
        for (int i = w.first_new(); i < w.cnt; ++i) {
          int slot;
          u32 par;
          w.slot_of(i, &slot, &par);
          mbar_wait_parked(&s_full[slot], par);
        }
        const u32 buf = t & (u32)(a.nbuf - 1), buf_ph = (t >> a.nbuf_sh) & 1u;
        mbar_wait_parked(&s_tempty[buf], buf_ph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          if (j < w.nr && !(ROWS_DBG(a) & 4)) {
            const u32 tmem_d = tmem_base + (buf * (u32)a.R + (u32)j) * a.acc_stride;
            int sl = w.sb + j * a.s;            // slot of this output row's first input row
            if (sl >= a.nslots) sl -= a.nslots;
            u32 rowoff = (u32)sl * slot16;
            const u32 wrap16 = (u32)a.nslots * slot16;
            u32 accumulate = 0;
            u32 ks = 0;
            for (int kh = 0; kh < d.KH; ++kh) {
              const u64 dbase = dslot0 + (u64)rowoff;
              for (int r = 0; r < steps_per_kh; ++r, ++ks) {
                const u64 da = dbase + (u64)a.aoff16[r];
                const u64 db = db0 + (u64)((ks >> 2) * bslab16 + 2u * (ks & 3u));
                umma_bf16(tmem_d, da, db, a.idesc, accumulate);
                accumulate = 1;
              }
              rowoff += slot16;                 // next filter row: one slot further
              if (rowoff >= wrap16) rowoff -= wrap16;
            }
          }
          // one commit per MMA warp and unit: s_tfull[buf] completes when every row's chain has retired.  The input
          // rows the unit frees are released by the epilogue (plain arrives) -- committing every freed slot from
          // every MMA warp cost ~50 clocks per tcgen05.commit, 72 of them per unit
          umma_commit(&s_tfull[buf]);
        }
        __syncwarp();
        w.next();
      }
    }
  } else {
    // =============================== epilogue ================================================================
    const int q = warp & 3;
    const int half = (warp - W_EPI0) >> 2;
    const int HoWo = d.Ho * d.Wo;
    int strip = u_begin / a.upr;
    int ug = u_begin - strip * a.upr;
    int n = strip / a.nseg, seg = strip - n * a.nseg;
    u32 t = 0;
    const bool releaser = warp == W_EPI0 && lane == 0;
    RowWalk rw;
    if (releaser) rw.start(u_begin, d.Ho, a.s, d.KH, a.R, a.upr, a.nslots);
    for (int u = u_begin; u < u_end; ++u, ++t) {
      const u32 buf = t & (u32)(a.nbuf - 1), buf_ph = (t >> a.nbuf_sh) & 1u;
      const int nr = min(a.R, d.Ho - ug * a.R);
      mbar_wait_parked(&s_tfull[buf], buf_ph);
      tc_fence_after();
      if (releaser) {   // every MMA of the unit has read its operands: free the input rows no later unit needs
        const int nrel = rw.released(u, u_end);
        int slot = rw.sb;
        for (int i = 0; i < nrel; ++i) {
          mbar_arrive(&s_empty[slot]);
          if (++slot == a.nslots) slot = 0;
        }
        rw.next();
      }
      const int ngroups = a.BN / 16;
      const int nitems = nr * ngroups;
      int j = 0, g = half;                 // item = (row j, group g), row-major; this warp takes every second one
      while (g >= ngroups) {
        g -= ngroups;
        ++j;
      }
      for (int it = half; it < nitems; it += 2) {
        const int opix = (ug * a.R + j) * d.Wo + seg * BM + 32 * q + lane;
        const int m = n * HoWo + opix;
        const u32 taddr = tmem_base + (buf * (u32)a.R + (u32)j) * a.acc_stride + ((u32)(32 * q) << 16);
        u32 v[16];
        tmem_ld16_nowait(taddr + (u32)(g * 16), v);
        tmem_ld_wait();
        const int co0 = g * 16;
        if (co0 < d.Co && !(ROWS_DBG(a) & 2)) epilogue_store(a, s_scale, s_shift, v, m, co0, co0, HoWo, n, opix);
        g += 2;
        while (g >= ngroups) {
          g -= ngroups;
          ++j;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[buf]);
      if (++ug == a.upr) {
        ug = 0;
        if (++seg == a.nseg) {
          seg = 0;
          ++n;
        }
      }
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, a.tmem_cols);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct RPlan {
  RArgs a;
  size_t smem;
};

// false when this geometry is not covered
static bool plan_rows(const cnb_conv_desc* d, RPlan* p) {
  static const bool off = [] { const char* e = getenv("CNB_CONV_ROWS"); return e && e[0] == '0'; }();
  if (off) return false;
  const int Ci = d->Ci, s = d->stride;
  if (!(Ci == 8 || Ci == 16 || Ci == 32 || Ci == 64)) return false;
  // filter KH x KW with vertical padding KH/2 and horizontal padding padx (= pad unless the descriptor says otherwise:
  // the space-to-depth stem is 7 x 5 with pad 3 / 2)
  const int padx = d->pad_w1 > 0 ? d->pad_w1 - 1 : d->pad;
  if (!(s == 1 || s == 2) || d->dil != 1 || d->KH % 2 == 0 || d->pad != d->KH / 2) return false;
  if (d->KW % 2 == 0 || padx != d->KW / 2) return false;
  if (d->KH < 3) return false;                       // 1x1: nothing to reuse, the im2col kernel is fine
  if (d->Wo % BM != 0 || d->Ho < 1) return false;
  if (d->Wi != d->Wo * s || d->Hi != d->Ho * s) return false;
  const int KWp = d->w_kw > 0 ? d->w_kw : d->KW;
  if (Ci == 8 && (s != 1 || KWp % 2 != 0 || KWp < d->KW)) return false;
  if (Ci != 8 && KWp != d->KW) return false;
  RArgs& a = p->a;
  a.s = s;
  a.nch = Ci / 8;
  a.KWp = KWp;
  // plane pixel range: q in [ox0 + floor(-pad/s), ox0 + 127 + floor((KWp-1-pad)/s)]
  a.q_off = -((padx + s - 1) / s);
  const int q_hi = (KWp - 1 - padx) / s;             // KWp-1-pad >= 0
  a.PW = BM + q_hi - a.q_off;
  int pw_alloc = a.PW;
  if (a.nch > 2)                                     // plane stride == 16 B mod 128 B: the 8 chunk planes of a pixel
    while ((pw_alloc * 16) % 128 != 16) ++pw_alloc;  // fall into distinct banks for the producers' stores
  a.plane_bytes = (u32)pw_alloc * 16u;
  a.slot_bytes = (u32)(s * a.nch) * a.plane_bytes;
  {
    const int items = s * a.nch * a.PW;
    a.npw = NPW;                           // rows are dealt out round-robin to the producer warps
    a.nk = (items + 31) / 32;
  }
  // ring: the KH rows of the current unit + the next unit's new rows + the rows in flight
  static const int env_depth = [] { const char* e = getenv("CNB_ROWS_DEPTH"); return e ? atoi(e) : 0; }();
  a.depth = env_depth > 0 ? env_depth : MAX_DEPTH;
  if (a.depth > MAX_DEPTH) a.depth = MAX_DEPTH;
  a.nseg = d->Wo / BM;
  a.BN = round_up(d->Co, 16);
  if (a.BN > 256) return false;
  const int Ktot = d->KH * KWp * Ci;
  const int Kpad = round_up(Ktot, 64);
  if (Ktot % 16 != 0 || Ktot / 16 > MAX_STEPS) return false;
  a.nslab = Kpad / 64;
  a.b_slab_bytes = (u32)a.BN * 128u;
  a.b_bytes = (u32)a.nslab * a.b_slab_bytes;
  a.b_bytes = (a.b_bytes + 1023u) & ~1023u;
  a.acc_stride = (u32)round_up(a.BN, 32);
  {
    const int ncp = a.nch >= 2 ? a.nch / 2 : 0;
    a.spk = ncp ? d->KW * ncp : KWp / 2;
    if (a.spk > 16) return false;
    for (int r = 0; r < a.spk; ++r) {
      u32 aoff;
      if (ncp) {
        const int kw = r / ncp, cp = r - kw * ncp;
        const int off = kw - padx;                       // input x = s*ox + kw - pad = s*(ox + dq) + phase
        const int dq = off >= 0 ? off / s : -((-off + s - 1) / s);
        const int ph = off - dq * s;
        aoff = (u32)(ph * a.nch + 2 * cp) * a.plane_bytes + (u32)(dq - a.q_off) * 16u;
      } else {
        aoff = (u32)(2 * r) * 16u;                       // 8 channels: one step = the taps kw = 2r, 2r+1
      }
      a.aoff16[r] = aoff >> 4;
    }
  }
  static const int env_r = [] { const char* e = getenv("CNB_ROWS_R"); return e ? atoi(e) : 0; }();
  static const int env_dbg = [] { const char* e = getenv("CNB_ROWS_DEBUG"); return e ? atoi(e) : 0; }();
  a.debug = env_dbg;
  a.idesc = make_idesc_bf16(BM, a.BN);
  // rows per unit: as many interleaved accumulators as TMEM (2 unit buffers) and the ring (shared memory) allow
  for (a.R = env_r > 0 ? (env_r > NMW ? NMW : env_r) : NMW;; --a.R) {
    if (a.R < 1) return false;
    if (2u * (u32)a.R * a.acc_stride > 512u) continue;
    a.nbuf = 2;
    while (a.nbuf < MAX_ACC && (u32)(2 * a.nbuf * a.R) * a.acc_stride <= 512u) a.nbuf *= 2;
    a.nbuf_sh = a.nbuf == 8 ? 3 : (a.nbuf == 4 ? 2 : 1);
    a.tmem_cols = 32;
    while (a.tmem_cols < (u32)(a.nbuf * a.R) * a.acc_stride) a.tmem_cols <<= 1;
    bool fits = false;
    for (int depth = a.depth; depth >= 2; --depth) {
      // ring: the unit's rows + the next unit's new rows + the rows in flight
      const int nslots = (a.R - 1) * s + d->KH + a.R * s + depth;
      const size_t smem = (size_t)a.b_bytes + (size_t)nslots * a.slot_bytes + (size_t)a.BN * 8 + 1024;
      if (nslots <= MAX_SLOTS && smem <= 224 * 1024) {
        a.nslots = nslots;
        a.depth = depth;
        p->smem = smem;
        fits = true;
        break;
      }
    }
    if (fits) break;
  }
  a.upr = (d->Ho + a.R - 1) / a.R;
  const long long units = (long long)d->B * a.nseg * a.upr;
  if (units >= (1ll << 31)) return false;
  a.units = (int)units;
  return true;
}

}  // namespace

bool conv_rows_supported(const cnb_conv_desc* d) {
  RPlan p;
  return plan_rows(d, &p);
}

int conv_rows_run(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale, const float* shift,
                  const void* res, void* y, cudaStream_t st) {
  TmaDriver& drv = tma_driver();
  if (!drv.ok) {
    set_error("conv: cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return CNB_ERR_CUDA;
  }
  RPlan p;
  CNB_CHECK_ARG(plan_rows(d, &p), "conv_rows: unsupported geometry");
  RArgs& a = p.a;
  a.d = *d;
  a.x = (const __nv_bfloat16*)x;
  a.scale = scale;
  a.shift = shift;
  a.res = (const __nv_bfloat16*)res;
  a.y = y;
  CUtensorMap tmB;
  {
    const int Kpad = a.nslab * 64;
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)a.BN};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)a.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = drv.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wpk, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_rows: cuTensorMapEncodeTiled failed (%d) Kpad=%d BN=%d", (int)r, Kpad, a.BN);
      return CNB_ERR_CUDA;
    }
  }
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(conv_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    once.mark();
  }
  const int grid = a.units < drv.num_sms ? a.units : drv.num_sms;
  CNB_CUDA(launch_pdl(conv_rows_kernel, dim3(grid), dim3(NTHREADS), p.smem, st, tmB, a));
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace cnb
