// Small band blocks: the log-likelihood with the Cholesky factor resident in shared memory.
//
// The reference's own data sets are small: N = 141 (ipynb/fitting_glQSO_example.ipynb), N = 194 in four band blocks of
// about 48 rows (ipynb/fitting_SLSN_example.ipynb), and with a lensing model the covariance -- so its factor -- is block
// diagonal by band (lensingmodels.py:39-51).  For such blocks (up to 224 rows) the factor of one band block
// fits the shared memory of an SM, and this kernel keeps it there: nothing of size N^2 touches global memory, the only
// HBM traffic of an evaluation is theta in and logL out.
//
// Replaces gaussn/gausSN.py:135-160 (GP.loglikelihood) exactly as gsn::logl_dmma_kernel does, with the same per-tile
// arithmetic in the same order -- results are bit-identical to that kernel (tests/test_gpu_onchip.py) -- but at 8-row
// granularity (N = 141 is factored as 144 rows, not 160) and without workspace, cp.async rings or item lists.
//
// Layout.  The rows of every band block start at a multiple of 8 (`OnchipDev`; padding rows are identity rows).  The factor
// is held as 8x8 tiles in the universal layout of gsn_chol_dmma.cuh (lane t owns the 16 bytes at offset 2t doubles:
// conflict-free 512-byte warp accesses, and a lane only ever touches its own 16 bytes of a tile).  The diagonal tile (j, j)
// holds -inv(L(j, j)): L(j, j) itself is never read again once its inverse exists.
// Only the LIVE part of the factor is stored.  A left-looking sweep reads row r for the last time in tile column r
// (as the B operand of that column), and nothing but logL leaves the kernel, so the tiles of row r are dead once every
// warp has finished column r, and tiles born two columns later take their slots.  The slot of tile (i, k) comes from a table
// built on the host by running exactly that schedule (onchip_slot_table): 100 slots instead of 171 at N = 144, which is four
// thetas per SM instead of two.  `cdone[j]` counts the warps that have finished column j; a warp enters column j + 2 (where
// it may overwrite row j) only when all have.
//
// Schedule.  A team of T warps (T = blockDim.x / 32 = 1, 2, 4, 8 or 16, chosen by the host so that an SM holds 16 warps)
// factors one theta; tile row i belongs to warp i mod T.  Left-looking by tile columns:
//     T(i, j), i > j:  S = sum_{k<j} L(i,k) L(j,k)^T - K(i,j);  L(i,j) = -S inv(L(j,j))^T         (2j + 2 DMMA)
//     D(j):            S = sum_{k<j} L(j,k) L(j,k)^T - K(j,j), 8x8 Cholesky + inverse by the warp; the residual b m(x~) - y
//                      rides along as row 0 of an extra tile Z, whose "L" is z^T = (L^-1 r)^T
// A warp sweeps the columns; in column j it first runs T(j+1, j) and D(j+1) back to back if row j+1 is its own
// (look-ahead: the next column's pivot tile is on the critical path), then its other rows two at a time (two independent
// accumulator chains per B operand; four at a time: 102 registers, -10 % at N = 64, -16 % at N = 141).  Row operands L(i, k) are the warp's own stores; the only cross-warp data are row j
// (needed from column j on) and -inv(L(j,j)), z(j).  Two monotone counters in shared memory carry that dataflow:
// prog_row (rows complete up to their last off-diagonal tile), prog_diag (pivot tiles factored) and prog_z (z stored),
// release / acquire at CTA scope; the GEMM of T(i, j) only needs prog_row > j, so it overlaps the 8x8 factorisation of D(j)
// and waits for prog_diag > j just before its solve; z is only read by the pivot rows, off the critical path.  Every wait is on work that precedes the waiter in column order: no deadlock.
// With T = 1 there is no flag traffic at all.
//
// logL = -1/2 (N ln 2pi + sum ln pivots + z^T z), reduced per 32-row group in a fixed order.
#pragma once
#include "gsn_chol_dmma.cuh"

namespace gsn {

#ifndef GSN_ONCHIP_DIAG_TILE
// Pivots by one shuffle from the tile's own diagonal (see chol8_inv_warp): ten DMMAs fewer per pivot tile, 18 cycles more per
// pivot.  With the register shape and five-warp teams the FP64 pipe, not the pivot chain, is the scarcer resource: N = 64
// 64.6 -> 71.0 M evals/s, N = 32 181 -> 204 M, the SN notebook's shape 21.6 -> 24.6 M, N = 141 11.06 -> 11.28 M (the one-warp
// shared-memory shape of early round 2 lost 18 % to it; the workspace kernels keep the broadcast diagonal tile: N = 512
// 525 k against 517 k).  The two families therefore differ in the last bits of a pivot (a[c][c] updated by fma against
// a[c][c] - sum round(L[c][k]^2)); tests hold them to 1e-9 + 1e-13 |logL| of each other.
#define GSN_ONCHIP_DIAG_TILE false
#ifndef GSN_ONCHIP_PIVOT_RANK2
#define GSN_ONCHIP_PIVOT_RANK2 0     // 1: two pivots per DMMA (chol8_inv_warp_r2: 10 instead of 18 DMMAs per pivot tile, 14 more shuffles per pair).
                                     // Correct (same tests green) but 3-5 % slower: N = 64 68.6 against 70.8 M evals/s, N = 32 194 against 204 M, N = 141 11.21
                                     // against 11.28 M -- at 18 DMMAs per tile the shapes are no longer bound by the FP64 pipe alone
#endif
#endif
#ifndef GSN_ONCHIP_LAG
#define GSN_ONCHIP_LAG 2
#endif
constexpr int kOnchipLag = GSN_ONCHIP_LAG;   // tiles written in column j take the slots of row j - kOnchipLag (teams of several warps)
// A one-warp team finishes column j - 1 (its last reads of row j - 1) before it stores anything of column j, so its tiles
// can take the slots of row j - 1 already: 21 slots instead of 25 at N = 64 (15 thetas per SM instead of 13), 7 instead of 9
// at N = 32.  The host builds both tables; the kernel picks by team size.
constexpr int kOnchipLagOneWarp = 1;
constexpr int kOnchipMaxTiles = 28;    // largest band block: 224 rows (225 live tiles = 113 KB; slot numbers fit a byte)
constexpr int kOnchipDispatchRows = 208;   // band blocks above this (one theta per SM) run the workspace kernels: see gsn_api.cu
constexpr int kOnchipMaxWarps = 16;
constexpr int kOnchipMaxBlocks = 256;

struct OnchipDev {
  int n_blocks;        // band blocks (1 without a band mask)
  int nt_max;          // tile rows of the largest block
  int n_groups;        // 32-row groups over all blocks: granularity of the final reduction
  int n_slots;         // tile slots of the largest block's live set (teams of several warps)
  int n_slots1;        // ... for a one-warp team (kOnchipLagOneWarp)
  const int4* blk;     // [n_blocks] {first row in the on-chip layout, tile rows, first group, real rows}
  const int* tab_off;  // [n_blocks] offset of the block's slot table in `tabs`
  const unsigned char* tabs;   // slot of tile (i, k) at i(i+1)/2 + k, per distinct block size
  const unsigned char* tabs1;  // the same for a one-warp team
  const double* x;     // per row of the on-chip layout (bands aligned to 8 rows)
  const double* y;
  const double* yerr2;
  const int* img;
  const int* urow;     // caller's row, -1 for padding rows
};

struct OnchipCtl {
  int fail;            // 0, or 1-based caller's row of the first non-positive pivot
  int next;            // theta index taken by the CTA
  int prog_row;        // rows [0, prog_row) are stored through their last off-diagonal tile
  int prog_diag;       // pivot tiles [0, prog_diag) are factored: -inv(L(j,j)) is stored
  int prog_z;          // z(0) ... z(prog_z - 1) are stored
  int cdone[kOnchipMaxTiles];   // warps that have finished tile column j
  double gmant[8];     // pivot product of each 32-row group of the block in work, as mantissa ...
  int gexpo[8];        // ... and exponent; the logarithms are taken after the factorisation, off its critical path
};

// Shared-memory carve-up (offsets in doubles).
struct OnchipSmem {
  int off_theta, off_ldg, off_zzg, off_ctl, off_vec, off_tab, off_tiles, vstride;
  size_t bytes;
  __host__ __device__ OnchipSmem(int nt_max, int n_slots, int n_groups, int P, bool has_aux) {
    int o = 64;                                   // 2^(j/64) table of fast_exp
    off_theta = o; o += (P + 1) & ~1;
    off_ldg = o; o += n_groups;
    off_zzg = o; o += n_groups;
    o = (o + 1) & ~1;
    off_ctl = o; o += (int)((sizeof(OnchipCtl) + 15) / 16 * 2);   // vectors and tiles are accessed 16 bytes at a time
    vstride = 8 * nt_max;
    off_vec = o; o += vstride * (has_aux ? 5 : 4);   // shifted epochs, magnifications, residual (overwritten by z), yerr^2, [aux]
    off_tab = o; o += (nt_max * (nt_max + 1) / 2 + 15) / 16 * 2;   // slot table of the block in work (bytes)
    off_tiles = o; o += n_slots * 64;
    bytes = (size_t)o * 8;
  }
};

// Slot table of a block of nt tile rows: tile (i, k) -> slot, by simulating the schedule.  Phase p (tile column p; p = -1
// is the first pivot tile) gives birth to the pivot tile (p + 1, p + 1) and the tiles (i, p), i > p; the slots of row
// p - lag are free again from phase p on.  Returns the number of slots.
inline int onchip_slot_table(int nt, unsigned char* tab, int lag = kOnchipLag) {
  auto tri = [](int i) { return i * (i + 1) / 2; };
  int free_list[kOnchipMaxTiles * (kOnchipMaxTiles + 1) / 2], head = 0, tail = 0, fresh = 0;
  auto take = [&]() { return head < tail ? free_list[head++] : fresh++; };
  for (int p = -1; p < nt - 1; ++p) {
    if (p >= lag)
      for (int k = 0; k <= p - lag; ++k) free_list[tail++] = tab[tri(p - lag) + k];
    tab[tri(p + 1) + p + 1] = (unsigned char)take();
    if (p >= 0)
      for (int i = p + 1; i < nt; ++i) tab[tri(i) + p] = (unsigned char)take();
  }
  return fresh;
}

__device__ __forceinline__ bool onchip_wait_gt(const int* p, int v, const int* fail) {
  while (ld_acquire_shared(p) <= v) {
    if (ld_volatile_shared(fail) != 0) return false;
    __nanosleep(20);
  }
  return true;
}

#ifdef GSN_ONCHIP_TRACE
// Debug build only: per-column timestamps of one theta (CTA 0, its GSN_ONCHIP_TRACE-th theta), read back by tools/.
__device__ long long g_onchip_trace[8192];
#define GSN_TRACE(slot, col) do { if (trace_on && lane == 0) g_onchip_trace[((col) + 1) * 16 + (slot)] = clock64(); } while (0)
#else
#define GSN_TRACE(slot, col) do { } while (0)
#endif

template <int KIND>
struct OnchipTeam {
  bool trace_on = false;
  KernelConsts kc;
  const double* exptab;
  const double *xs, *bs, *rs, *e2, *aux;
  double* zs;
  double2* tiles;      // already offset by lane
  const unsigned char* tab;   // slot of tile (i, k) at tri(i) + k
  OnchipCtl* ctl;
  double* ldg;
  const int* urow;     // of this block
  int lane, T, nt, nrows, n_total;

  struct Cols { double2 x, ba, a; };   // per-lane column data of a tile column: epochs, magnification * amplitude, auxiliary

  __device__ __forceinline__ Cols load_cols(int j) const {
    const int c = 8 * j + (lane & 3) * 2;
    Cols q;
    q.x = *reinterpret_cast<const double2*>(xs + c);
    const double2 b = *reinterpret_cast<const double2*>(bs + c);
    q.ba = make_double2(b.x * kc.amp, b.y * kc.amp);      // gausSN.py:146 with the kernel's leading factor folded in
    q.a = kernel_has_aux<KIND>() ? *reinterpret_cast<const double2*>(aux + c) : make_double2(0.0, 0.0);
    return q;
  }

  // -K(i, j) for one tile: shifted epochs, magnifications, noise diagonal (gausSN.py:146-147), identity on padding rows.
  template <bool kBranch = false>
  __device__ __forceinline__ double2 build(int i, int j, const Cols& q) const {
    const int row = 8 * i + (lane >> 2), col = 8 * j + (lane & 3) * 2;
    const double xr = xs[row], nbr = -bs[row];
    const double ar = kernel_has_aux<KIND>() ? aux[row] : 0.0;
    double2 v;
    v.x = (nbr * q.ba.x) * cov_unit_sym<KIND, kBranch>(kc, xr, q.x.x, ar, q.a.x, exptab);
    v.y = (nbr * q.ba.y) * cov_unit_sym<KIND, kBranch>(kc, xr, q.x.y, ar, q.a.y, exptab);
    if (i == j) {
      if (row == col) v.x -= e2[row];
      if (row == col + 1) v.y -= e2[row];
    }
    if (i == nt - 1) {   // the block's last tile row may carry padding rows (and, on the diagonal tile, padding columns)
      if (row >= nrows || col >= nrows) v.x = (row == col) ? -1.0 : 0.0;
      if (row >= nrows || col + 1 >= nrows) v.y = (row == col + 1) ? -1.0 : 0.0;
    }
    return v;
  }

  __device__ __forceinline__ static int tri(int i) { return i * (i + 1) / 2; }
  __device__ __forceinline__ double2* tile(const unsigned char* row, int k) const { return tiles + (int)row[k] * 32; }
  // Residual row.  y(i) = sum_k L(i,k) z(k) is accumulated as one partial sum per lane (its two columns of every tile, two
  // DFMAs per tile: an 8x8 matrix-vector product issued as DMMAs would occupy the FP64 pipe eight times as long), reduced over
  // the four lanes of a row at the end; z(i) = inv(L(i,i)) (r(i) - y(i)).
  __device__ __forceinline__ double2 zpair(int k) const { return *reinterpret_cast<const double2*>(zs + 8 * k + 2 * (lane & 3)); }
  __device__ __forceinline__ static double ydot(double y, const double2& a, const double2& z) { return fma(a.y, z.y, fma(a.x, z.x, y)); }
  // returns z(i) of this lane's row; stores it
  __device__ __forceinline__ double finish_z(int i, double y, const double2& nW) const {
    constexpr unsigned kFull = 0xffffffffu;
    const int r = lane >> 2, q = lane & 3;
    y += __shfl_xor_sync(kFull, y, 1);
    y += __shfl_xor_sync(kFull, y, 2);
    const double t = y - rs[8 * i + r];                       // -(r - L z), row r
    const double t0 = __shfl_sync(kFull, t, 8 * q), t1 = __shfl_sync(kFull, t, 8 * q + 4);   // rows 2q, 2q + 1
    double p = fma(nW.y, t1, nW.x * t0);                      // nW = -inv(L(i,i))
    p += __shfl_xor_sync(kFull, p, 1);
    p += __shfl_xor_sync(kFull, p, 2);
    if (q == 0) zs[8 * i + r] = p;
    return p;
  }

  // Before the first tile store of column j: the slots it may take are those of row j - kOnchipLag, dead once every warp has
  // left that column.
  __device__ __forceinline__ bool may_write(int j, bool& write_ok) const {
    if (T > 1 && !write_ok) {
      if (j >= kOnchipLag && !onchip_wait_gt(&ctl->cdone[j - kOnchipLag], T - 1, &ctl->fail)) return false;
      write_ok = true;
    }
    return true;
  }

  // Column j for R of this warp's rows: T(i[r], j).  Returns false if the theta failed while waiting.
  template <int R>
  __device__ __forceinline__ bool panel(const int (&i)[R], int j, const Cols& q, bool& row_ok, bool& diag_ok, bool& write_ok) const {
    double2 S[R];
#pragma unroll
    for (int r = 0; r < R; ++r) S[r] = build(i[r], j, q);
    if (T > 1 && !row_ok) {
      if (!onchip_wait_gt(&ctl->prog_row, j, &ctl->fail)) return false;
      row_ok = true;
    }
    const unsigned char* tb = tab + tri(j);
    const unsigned char* ta[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ta[r] = tab + tri(i[r]);
#pragma unroll 2
    for (int k = 0; k < j; ++k) {
      const double2 b = *tile(tb, k);
      double2 a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = *tile(ta[r], k);
#pragma unroll
      for (int r = 0; r < R; ++r) dmma(S[r], a[r].x, b.x);
#pragma unroll
      for (int r = 0; r < R; ++r) dmma(S[r], a[r].y, b.y);
    }
    if (T > 1 && !diag_ok) {
      if (!onchip_wait_gt(&ctl->prog_diag, j, &ctl->fail)) return false;
      diag_ok = true;
    }
    const double2 nW = *tile(tb, j);
    if (!may_write(j, write_ok)) return false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double2 X = make_double2(0.0, 0.0);
      dmma(X, S[r].x, nW.x);
      dmma(X, S[r].y, nW.y);
      *tile(ta[r], j) = X;
    }
    return true;
  }

  // Look-ahead: T(i, j) for i = j + 1 (skipped for j = -1) and D(i) back to back.
  // Critical path of the whole factorisation: prog_diag > j  ->  solve  ->  D += X X^T  ->  8x8 elimination  ->  prog_diag = i + 1.
  // Everything else (the GEMMs over k < j, the residual row Z, the pivot products) is kept off it.
  __device__ __forceinline__ bool pivot_row(int j, bool& write_ok) const {
    const int i = j + 1;
    GSN_TRACE(0, j);
    const Cols qi = load_cols(i);
    double2 S = make_double2(0.0, 0.0), X = make_double2(0.0, 0.0);
    double2 D = build(i, i, qi);
    double y = 0.0;
    const unsigned char* ta = tab + tri(i);
    if (j >= 0) {
      S = build(i, j, load_cols(j));
      GSN_TRACE(1, j);
      if (T > 1 && !(onchip_wait_gt(&ctl->prog_row, j, &ctl->fail) && onchip_wait_gt(&ctl->prog_z, j - 1, &ctl->fail))) return false;
      GSN_TRACE(2, j);
      const unsigned char* tb = tab + tri(j);
#pragma unroll 2
      for (int k = 0; k < j; ++k) {
        const double2 b = *tile(tb, k);
        const double2 a = *tile(ta, k);
        const double2 z = zpair(k);
        dmma(S, a.x, b.x);
        dmma(D, a.x, a.x);
        dmma(S, a.y, b.y);
        dmma(D, a.y, a.y);
        y = ydot(y, a, z);
      }
      GSN_TRACE(3, j);
      if (T > 1 && !onchip_wait_gt(&ctl->prog_diag, j, &ctl->fail)) return false;
      GSN_TRACE(4, j);
      const double2 nWj = *tile(tb, j);
      dmma(X, S.x, nWj.x);
      dmma(X, S.y, nWj.y);
      if (!may_write(j, write_ok)) return false;
      *tile(ta, j) = X;
      if (T > 1) {
        __syncwarp();
        if (lane == 0) st_release_shared(&ctl->prog_row, i + 1);
      }
      dmma(D, X.x, X.x);
      dmma(D, X.y, X.y);
    }
    double mant = 1.0;
    int expo = 0, failrow = -1;
    double2 nW;
    GSN_TRACE(5, j);
#if GSN_ONCHIP_PIVOT_RANK2
    chol8_inv_warp_r2(make_double2(-D.x, -D.y), lane, nW, mant, expo, failrow);
#else
    chol8_inv_warp<GSN_ONCHIP_DIAG_TILE>(make_double2(-D.x, -D.y), lane, nW, mant, expo, failrow);
#endif
    *tile(ta, i) = nW;
    if (T > 1) {
      __syncwarp();
      if (lane == 0) st_release_shared(&ctl->prog_diag, i + 1);
    }
    GSN_TRACE(6, j);
    // ---- off the critical path: the residual row, the pivot product, the failure index
    if (j >= 0) {
      if (T > 1 && !onchip_wait_gt(&ctl->prog_z, j, &ctl->fail)) return false;
      y = ydot(y, X, zpair(j));
    }
    finish_z(i, y, nW);
    if (lane == 0) {
      if (failrow >= 0) {   // a padding row can only get here through NaNs that an earlier row produced
        const int u = urow[8 * i + failrow];
        atomicCAS(&ctl->fail, 0, u >= 0 ? u + 1 : n_total);
      }
      // running pivot product of the 32-row group: continues where D(i - 1) left it (pivot rows run in order)
      if ((i & 3) != 0) { mant *= ctl->gmant[i >> 2]; expo += ctl->gexpo[i >> 2]; }
      ctl->gmant[i >> 2] = mant;
      ctl->gexpo[i >> 2] = expo;
    }
    __syncwarp();
    if (T > 1 && lane == 0) st_release_shared(&ctl->prog_z, i + 1);
    GSN_TRACE(7, j);
    return true;
  }

  __device__ __forceinline__ void factor(int warp) const {
    int first = warp;          // this warp's first row >= j + 1, kept incrementally (as (j + 1) + (warp - (j + 1) % T + T) % T the
                               // two divisions by a run-time T were 9 % of all instructions executed at N = 141)
    for (int j = -1; j < nt - 1; ++j) {
      if (ld_volatile_shared(&ctl->fail) != 0) break;
      if (first < j + 1) first += T;
      // Rows are dealt out cyclically.  (The work of a row grows like i^2, so the warp that gets the last row carries 1.3 x the
      // mean at 18 rows on 5 warps; dealing the rows out from the bottom in a snake brings that to 1.07 x -- measured: N = 141
      // 11.37 against 11.28 M evals/s, three bands of 132 rows 3.55 against 4.16 M, and 9.4 M at N = 141 when a warp then takes its
      // rows nearest to the diagonal first: the warps' relative timing matters more than their total load.  Likewise letting the
      // owner of the NEXT pivot row run it early -- its tile of this column first, then pivot_row(j + 1), then its other rows; the
      // per-column trace shows the chain standing still for 4-6 k cycles where that owner is late -- costs 20 % (N = 141 9.0 M):
      // the warp then idles in the pivot row's waits instead of working on its panels, and with four thetas per SM it is the
      // busy time, not one theta's critical path, that counts.)
      int i = first;
      bool ok = true, write_ok = false;
      if (i == j + 1 && i < nt) {
        ok = pivot_row(j, write_ok);
        i += T;
      }
      if (ok && j >= 0 && i < nt) {
        const Cols q = load_cols(j);
        bool row_ok = false, diag_ok = false;
        for (; ok && i + T < nt; i += 2 * T) {
          const int ii[2] = {i, i + T};
          ok = panel<2>(ii, j, q, row_ok, diag_ok, write_ok);
        }
        if (ok && i < nt) {
          const int ii[1] = {i};
          ok = panel<1>(ii, j, q, row_ok, diag_ok, write_ok);
        }
      }
      if (!ok) break;
      if (T > 1 && j >= 0) {
        __syncwarp();
        if (lane == 0) {
          asm volatile("red.release.cta.shared.add.s32 [%0], 1;" :: "r"(smem_u32(&ctl->cdone[j])) : "memory");
        }
      }
    }
  }
};

// Band blocks of at most kRegMaxTiles tile rows (64 rows: the getting-started shape, the SN notebook's four bands of 48-49
// rows) factored by ONE warp with the factor in REGISTERS.  The universal tile layout lets an accumulator tile be fed back
// as an operand as it is, so the left-looking sweep of OnchipTeam needs no storage at all once its loops are unrolled over a
// compile-time block size: tile (i, k) is a register pair per lane, the compiler's liveness analysis does what the slot
// table does for the shared-memory shape, and every load, store, slot lookup and loop counter of the sweep disappears
// (N = 64: 9 500 instructions per theta through shared memory; 56 -> 65 M evals/s, the SN notebook's shape 20 -> 21.6 M).  All rows of a tile column are swept together (up to six
// independent accumulator chains, and as many exponentials in flight in the build).  The arithmetic per tile and its
// order are those of OnchipTeam::pivot_row / panel: results are bit-identical (tests/test_gpu_onchip.py).
constexpr int kRegMaxTiles = 8;
#ifndef GSN_REG_CTAS
#define GSN_REG_CTAS 16            // one-warp CTAs per SM the register shape is compiled for: 128 registers, 8 bytes of spills
                                   // (12 CTAs / 154 registers: N = 32 167 M against 181 M evals/s; 20 CTAs / 96 registers: 123 M)
#endif
constexpr int kRegCtasPerSm = GSN_REG_CTAS;

#ifndef GSN_REG_CALL_CHOL
#define GSN_REG_CALL_CHOL 1        // the 8x8 pivot routine as one out-of-line copy (the unrolled sweep would inline it up to 36 times)
#endif
#ifndef GSN_REG_CALL_BUILD
#define GSN_REG_CALL_BUILD 0       // ... but not the covariance build of a tile (the call costs more than the instruction cache misses: N = 64 52 M
                                   // against 64 M evals/s; one out-of-line loop per tile column, results handed over through shared memory: 47 M)
#endif
static __device__ __noinline__ Chol8Out chol8_inv_onchip_call(double2 a, int lane) {
  Chol8Out o;
  o.mant = 1.0; o.expo = 0; o.failrow = -1;
#if GSN_ONCHIP_PIVOT_RANK2
  chol8_inv_warp_r2(a, lane, o.nW, o.mant, o.expo, o.failrow);
#else
  chol8_inv_warp<GSN_ONCHIP_DIAG_TILE>(a, lane, o.nW, o.mant, o.expo, o.failrow);
#endif
  return o;
}
template <int KIND>
static __device__ __noinline__ double2 onchip_build_call(const OnchipTeam<KIND>* tm, int i, int j) {
  return tm->template build<true>(i, j, tm->load_cols(j));
}
template <int KIND>
__device__ __forceinline__ double2 reg_build(const OnchipTeam<KIND>& tm, int i, int j, const typename OnchipTeam<KIND>::Cols& q) {
#if GSN_REG_CALL_BUILD
  return onchip_build_call<KIND>(&tm, i, j);
#else
  return tm.template build<true>(i, j, q);
#endif
}

template <int KIND, int NT>
__device__ __forceinline__ void reg_factor(const OnchipTeam<KIND>& tm) {
  using Cols = typename OnchipTeam<KIND>::Cols;
  const int lane = tm.lane;
  double2 L[NT][NT];   // L[i][k], k < i: the factor's tile; L[i][i]: -inv(L(i, i))
#pragma unroll
  for (int j = -1; j < NT - 1; ++j) {
    const int i = j + 1;
    // ---- T(j + 1, j) and D(j + 1): OnchipTeam::pivot_row
    {
      const Cols qi = tm.load_cols(i);
      double2 S = make_double2(0.0, 0.0), X = make_double2(0.0, 0.0);
      double2 D = reg_build<KIND>(tm, i, i, qi);
      double y = 0.0;
      if (j >= 0) {
        S = reg_build<KIND>(tm, i, j, tm.load_cols(j));
#pragma unroll
        for (int k = 0; k < j; ++k) {
          const double2 a = L[i][k], b = L[j][k];
          const double2 z = tm.zpair(k);
          dmma(S, a.x, b.x);
          dmma(D, a.x, a.x);
          dmma(S, a.y, b.y);
          dmma(D, a.y, a.y);
          y = OnchipTeam<KIND>::ydot(y, a, z);
        }
        const double2 nWj = L[j][j];
        dmma(X, S.x, nWj.x);
        dmma(X, S.y, nWj.y);
        L[i][j] = X;
        dmma(D, X.x, X.x);
        dmma(D, X.y, X.y);
      }
      double mant = 1.0;
      int expo = 0, failrow = -1;
      double2 nW;
#if GSN_REG_CALL_CHOL
      {
        const Chol8Out o = chol8_inv_onchip_call(make_double2(-D.x, -D.y), lane);
        nW = o.nW; mant = o.mant; expo = o.expo; failrow = o.failrow;
      }
#elif GSN_ONCHIP_PIVOT_RANK2
      chol8_inv_warp_r2(make_double2(-D.x, -D.y), lane, nW, mant, expo, failrow);
#else
      chol8_inv_warp<GSN_ONCHIP_DIAG_TILE>(make_double2(-D.x, -D.y), lane, nW, mant, expo, failrow);
#endif
      L[i][i] = nW;
      if (j >= 0) y = OnchipTeam<KIND>::ydot(y, X, tm.zpair(j));
      tm.finish_z(i, y, nW);
      if (lane == 0) {
        if (failrow >= 0) {
          const int u = tm.urow[8 * i + failrow];
          atomicCAS(&tm.ctl->fail, 0, u >= 0 ? u + 1 : tm.n_total);
        }
        if ((i & 3) != 0) { mant *= tm.ctl->gmant[i >> 2]; expo += tm.ctl->gexpo[i >> 2]; }
        tm.ctl->gmant[i >> 2] = mant;
        tm.ctl->gexpo[i >> 2] = expo;
      }
      __syncwarp();
    }
    // ---- the other rows of tile column j, all at once: OnchipTeam::panel
    if (j >= 0 && j + 2 < NT) {
      const Cols q = tm.load_cols(j);
      double2 S[NT];
#pragma unroll
      for (int r = j + 2; r < NT; ++r) S[r] = reg_build<KIND>(tm, r, j, q);
#pragma unroll
      for (int k = 0; k < j; ++k) {
        const double2 b = L[j][k];
#pragma unroll
        for (int r = j + 2; r < NT; ++r) dmma(S[r], L[r][k].x, b.x);
#pragma unroll
        for (int r = j + 2; r < NT; ++r) dmma(S[r], L[r][k].y, b.y);
      }
      const double2 nW = L[j][j];
#pragma unroll
      for (int r = j + 2; r < NT; ++r) {
        double2 X = make_double2(0.0, 0.0);
        dmma(X, S[r].x, nW.x);
        dmma(X, S[r].y, nW.y);
        L[r][j] = X;
      }
    }
  }
}

#ifndef GSN_REG_MAXNT
#define GSN_REG_MAXNT 8
#endif
template <int KIND>
__device__ __forceinline__ void reg_factor_any(const OnchipTeam<KIND>& tm, int nt) {
  switch (nt) {
    case 1: reg_factor<KIND, 1>(tm); break;
    case 2: reg_factor<KIND, 2>(tm); break;
    case 3: reg_factor<KIND, 3>(tm); break;
    case 4: reg_factor<KIND, 4>(tm); break;
#if GSN_REG_MAXNT >= 6
    case 5: reg_factor<KIND, 5>(tm); break;
    case 6: reg_factor<KIND, 6>(tm); break;
#endif
#if GSN_REG_MAXNT >= 7
    case 7: reg_factor<KIND, 7>(tm); break;
#endif
#if GSN_REG_MAXNT >= 8
    default: reg_factor<KIND, 8>(tm); break;
#endif
  }
}

template <int KIND, bool kReg>
#ifndef GSN_ONCHIP_BOUND_THREADS
#define GSN_ONCHIP_BOUND_THREADS 640   // more than the 512 threads ever launched: caps the shared-memory shape at 96 registers (21 warps per SM)
#endif
__global__ void __launch_bounds__(kReg ? 32 : GSN_ONCHIP_BOUND_THREADS, kReg ? kRegCtasPerSm : 1)
logl_onchip_kernel(ProblemDev pd, OnchipDev od, const double* __restrict__ theta, long long B, long long ld,
                   const double* __restrict__ hostmean, GatherOut out, int* __restrict__ info, unsigned flags,
                   unsigned* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const bool one_warp = (nthreads == 32);
  const int n_slots = kReg ? 0 : (one_warp ? od.n_slots1 : od.n_slots);
  const OnchipSmem lay(od.nt_max, n_slots, od.n_groups, pd.P, kernel_has_aux<KIND>());
  double* exptab = sm;
  double* th = sm + lay.off_theta;
  double* ldg = sm + lay.off_ldg;
  double* zzg = sm + lay.off_zzg;
  OnchipCtl* ctl = reinterpret_cast<OnchipCtl*>(sm + lay.off_ctl);
  double* xs = sm + lay.off_vec;
  double* bs = xs + lay.vstride;
  double* rs = bs + lay.vstride;
  double* e2 = rs + lay.vstride;
  double* zs = rs;      // z(i) replaces the residual of tile row i (read once, by the warp that then writes z(i))
  double* aux = kernel_has_aux<KIND>() ? e2 + lay.vstride : xs;

  for (int i = tid; i < 64; i += nthreads) exptab[i] = kExp2Tab[i];
#ifdef GSN_ONCHIP_TRACE
  int n_done = -1;
#endif

  for (;;) {
#ifdef GSN_ONCHIP_TRACE
    ++n_done;
#endif
    __syncthreads();
    if (tid == 0) ctl->next = (int)atomicAdd(counter, 1u);
    __syncthreads();
    const long long it = ctl->next;
    if (it >= B) break;

    // ---- theta (gausSN.py:173-194: kernel | mean | lensing, fixed slots from the handle)
    for (int k = tid; k < pd.P; k += nthreads) {
      const int col = pd.theta_col[k];
      th[k] = (col >= 0) ? theta[it * ld + col] : pd.theta_fixed[k];
    }
    if (tid == 0) ctl->fail = 0;
    __syncthreads();
    const double* mp = th + pd.nk;
    const double* lp = th + pd.nk + pd.nm;
    OnchipTeam<KIND> team;
    team.kc = kernel_consts<KIND>(th);
    team.exptab = exptab; team.xs = xs; team.bs = bs; team.rs = rs; team.e2 = e2; team.aux = aux; team.zs = zs;
    team.tiles = reinterpret_cast<double2*>(sm + lay.off_tiles) + lane;
    unsigned char* tab = reinterpret_cast<unsigned char*>(sm + lay.off_tab);
    team.tab = tab;
    team.ctl = ctl; team.ldg = ldg; team.lane = lane; team.T = nthreads >> 5; team.n_total = pd.N;
    const KernelConsts& kc = team.kc;
    // a non-finite kernel constant or shifted epoch makes the covariance NaN, which the reference's Cholesky turns into
    // NaN: detected here, once per theta, so that the N^2 build can use the branch-free exp core
    int bad = (tid == 0) && !(isfinite(kc.amp) && isfinite(kc.c1) && isfinite(kc.c2) && isfinite(kc.c3) && isfinite(kc.c4));
    bool dead = false;

    for (int blk = 0; blk < od.n_blocks; ++blk) {
      const int4 bd = od.blk[blk];
      const int r0 = bd.x, nt = bd.y, g0 = bd.z, nrows = bd.w;
      // ---- per-epoch vectors of the block: shifted epochs, magnifications, residual (gausSN.py:139-142)
      for (int n = tid; n < 8 * nt; n += nthreads) {
        double x_s = 0.0, b = 0.0, r = 0.0, ax = 0.0, e = 1.0;
        const int u = od.urow[r0 + n];
        if (u >= 0) {
          lens_eval(pd.lens_kind, lp, od.img[r0 + n], od.x[r0 + n], x_s, b);
          const double m = hostmean ? hostmean[it * pd.N + u] : mean_eval(pd.mean_kind, mp, x_s);
          r = b * m - od.y[r0 + n];
          e = od.yerr2[r0 + n];
          if (kernel_has_aux<KIND>()) ax = kernel_aux<KIND>(kc, x_s);
          bad |= !isfinite(x_s) || !isfinite(ax);
        }
        xs[n] = x_s; bs[n] = b; rs[n] = r; e2[n] = e;
        if (kernel_has_aux<KIND>()) aux[n] = ax;
      }
      if constexpr (!kReg) {
        const unsigned char* src = (one_warp ? od.tabs1 : od.tabs) + od.tab_off[blk];
        for (int n = tid; n < nt * (nt + 1) / 2; n += nthreads) tab[n] = src[n];
        for (int n = tid; n < nt; n += nthreads) ctl->cdone[n] = 0;
      }
#ifdef GSN_ONCHIP_POISON
      // Debug build (compute-sanitizer is closed on this pool): every tile slot starts as NaN, so a tile that is read before
      // it was written in THIS block -- a wrong slot table, a missing wait -- turns the result into NaN and fails the parity
      // tests.  Rows whose slots are recycled are poisoned again by whoever recycles them (see may_write).
      for (int n = tid; n < n_slots * 32; n += nthreads)
        reinterpret_cast<double2*>(sm + lay.off_tiles)[n] = make_double2(__longlong_as_double(0x7ff8dead00000000LL), __longlong_as_double(0x7ff8dead00000000LL));
#endif
      if (tid == 0) { ctl->prog_row = 1; ctl->prog_diag = 0; ctl->prog_z = 0; }
      if (__syncthreads_or(bad)) { dead = true; break; }
      bad = 0;
      team.nt = nt; team.nrows = nrows; team.urow = od.urow + r0;
#ifdef GSN_ONCHIP_TRACE
      team.trace_on = (blockIdx.x == 0 && blk == 0 && n_done == GSN_ONCHIP_TRACE);
      if (team.trace_on && tid == 0) { g_onchip_trace[0] = clock64(); g_onchip_trace[1] = nt; }
#endif
      if constexpr (kReg) reg_factor_any<KIND>(team, nt); else team.factor(warp);
      __syncthreads();
#ifdef GSN_ONCHIP_TRACE
      if (team.trace_on && tid == 0) g_onchip_trace[2] = clock64();
#endif
      if (ctl->fail != 0) break;
      // z^T z per 32-row group: lane 4 g + q sums the squares of its two columns over the group's tiles
      if (warp == 0) {
        const int g = lane >> 2, q = lane & 3;
        double ss = 0.0;
        for (int b = 0; b < 4; ++b) {
          const int t = 4 * g + b;
          if (t < nt) {
            const double2 z2 = *reinterpret_cast<const double2*>(zs + 8 * t + 2 * q);
            ss = fma(z2.x, z2.x, ss);
            ss = fma(z2.y, z2.y, ss);
          }
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        if (q == 0 && 4 * g < nt) {
          zzg[g0 + g] = ss;
          ldg[g0 + g] = log(ctl->gmant[g]) + (double)ctl->gexpo[g] * 0.6931471805599453094;
        }
      }
    }

    // ---- logL = -1/2 (N ln 2 pi + ln det + z^T z)      gausSN.py:151-158
    __syncthreads();
    if (tid == 0) {
      double ll;
      int inf = 0;
      if (dead) { ll = nan(""); inf = -1; }
      else if (ctl->fail != 0) { ll = nan(""); inf = ctl->fail; }
      else {
        double zsum = 0.0, ldsum = 0.0;
        for (int g = 0; g < od.n_groups; ++g) { zsum += zzg[g]; ldsum += ldg[g]; }   // fixed order: run-to-run identical
        ll = -0.5 * (pd.nlog2pi + ldsum + zsum);
        if (!isfinite(ll)) inf = -1;
      }
      if ((flags & GSN_FLAG_NEG_INF) && !isfinite(ll)) ll = -INFINITY;
      for (int r = 0; r < out.n; ++r) out.ptr[r][out.offset + it] = ll;
      if (info) info[it] = inf;
    }
  }
}

}  // namespace gsn
