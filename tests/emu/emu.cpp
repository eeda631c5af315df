// TEST-ONLY: single-thread host build of the kernel source (csrc/mpc_core.h).
//
// It exists so that the algorithm the CUDA kernel runs (closed-form assembly,
// sweep inversion, dual active-set iterations) can be logic-checked against the
// oracle in the CPU test tier, where there is no GPU.  It is never linked into
// libquadruped_mpc_b200.so and nothing in quadruped_ctrl_b200/ loads it: the
// product has no CPU path.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../quadruped_ctrl_b200/csrc/mpc_core.h"
#include "../../quadruped_ctrl_b200/csrc/mpc_ticks.h"
#include "../../quadruped_ctrl_b200/csrc/mpc_legs.h"
#include "../../quadruped_ctrl_b200/csrc/mpc_riccati.h"

static std::vector<double> g_dbg_u, g_dbg_minv;
static std::vector<int> g_dbg_W;
static int g_dbg_nv = 0, g_dbg_m = 0;

// Runs `batch` records through the kernel body with one emulated thread.
//   nv_cap / m_cap <= 0 -> worst case (12h).  H_out/g_out (optional): the reduced QP
//   before inversion, [batch*(12h)^2] / [batch*12h] with leading dimension 12h.
//   info [batch*4]: nv, active-set size at exit, iterations, status code.
template <bool PK>
static int emu_solve_batch_t(const void* records, int batch, int h, int nv_cap, int m_cap, int max_iter, float* forces,
                             double* solution, int* info, double* H_out, double* g_out, int* warm_cache = nullptr,
                             int warm_shift = 0) {
  using namespace mpc;
  if (nv_cap <= 0) nv_cap = 12 * h;
  if (m_cap <= 0) m_cap = nv_cap;
  const int packed = PK ? 1 : 0;
  const Layout L = make_layout(h, nv_cap, m_cap, 1, 0, packed);
  std::vector<char> fast(L.fast_bytes + 64);
  const size_t stride = ((size_t)(4 * (MPC_REC_TRAJ + 12 * h) + 4 * h) + 15) / 16 * 16;
  OneThreadT<PK> cx{0, 1};
  const int NU = 12 * h;
  for (int b = 0; b < batch; b++) {
    memset(fast.data(), 0xCD, fast.size());  // poison: nothing may rely on zeroed workspace
    const Work k = carve(L, fast.data(), nullptr);
    const float* rec = (const float*)((const char*)records + stride * b);
    const unsigned char* gait = (const unsigned char*)rec + 4 * (MPC_REC_TRAJ + 12 * h);
    assemble(cx, rec, gait, k);
    if (k.sc->status == MPC_STATUS_OPTIMAL && k.sc->nv > nv_cap) return -1;
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      if (H_out)
        for (int i = 0; i < k.sc->nv; i++)
          for (int j = 0; j < k.sc->nv; j++) H_out[(size_t)b * NU * NU + (size_t)i * NU + j] = k.Hm[hix(k.ld, i, j)];
      if (g_out)
        for (int i = 0; i < k.sc->nv; i++) g_out[(size_t)b * NU + i] = k.g[i];
      if (k.ld > 0) {
        invert_spd(cx, k);
      } else {
        // packed layout: the device inverts in registers (invert_spd_tiles); here the same symmetric sweep runs on
        // a dense copy, so that the packed stores of the assembly and the packed reads of the active set are tested
        const int n = k.sc->nv;
        std::vector<double> D((size_t)n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) D[(size_t)i * n + j] = k.Hm[hix(k.ld, i, j)];
        for (int p = 0; p < n && k.sc->status == MPC_STATUS_OPTIMAL; p++) {
          const double d = D[(size_t)p * n + p];
          if (!(d > 0.0)) { k.sc->status = MPC_STATUS_NOT_PD; break; }
          std::vector<double> c(n);
          for (int i = 0; i < n; i++) c[i] = D[(size_t)p * n + i];
          for (int i = 0; i < n; i++) for (int j = 0; j < n; j++)
            if (i != p && j != p) D[(size_t)i * n + j] -= c[i] * c[j] / d;
          for (int i = 0; i < n; i++) { D[(size_t)p * n + i] = c[i] / d; D[(size_t)i * n + p] = c[i] / d; }
          D[(size_t)p * n + p] = -1.0 / d;
        }
        for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) k.Hm[hix(k.ld, i, j)] = -D[(size_t)i * n + j];
      }
      if (k.sc->status == MPC_STATUS_OPTIMAL) {
        active_set_init(cx, rec, gait, k);
        if (warm_cache) active_set_warm(cx, rec, k, warm_cache + (size_t)b * kWarmStride, warm_shift);
        active_set(cx, rec, gait, k, max_iter);
      }
    }
    g_dbg_m = k.sc->m; g_dbg_nv = k.sc->nv;
    g_dbg_W.assign(k.W, k.W + g_dbg_m); g_dbg_u.assign(k.u, k.u + g_dbg_m);
    g_dbg_minv.resize((size_t)g_dbg_nv * g_dbg_nv);
    for (int i = 0; i < g_dbg_nv; i++) for (int j = 0; j < g_dbg_nv; j++) g_dbg_minv[(size_t)i * g_dbg_nv + j] = k.Hm[hix(k.ld, i, j)];
    if (warm_cache) {
      if (k.sc->status != MPC_STATUS_OPTIMAL) k.sc->m = 0;
      active_set_store(cx, k, warm_cache + (size_t)b * kWarmStride);
    }
    int32_t st = 0;
    scatter(cx, k, forces + 12 * b, solution ? solution + (size_t)NU * b : nullptr, &st);
    if (info) {
      info[4 * b + 0] = k.sc->nv;
      info[4 * b + 1] = k.sc->m;
      info[4 * b + 2] = k.sc->iters;
      info[4 * b + 3] = k.sc->status;
    }
  }
  return 0;
}

extern "C" {

// debugging aid: working set, duals and inverse of the LAST problem solved
int emu_debug_last(int* W, double* u, double* minv, int* nv) {
  for (int i = 0; i < g_dbg_m; i++) { W[i] = g_dbg_W[i]; u[i] = g_dbg_u[i]; }
  for (size_t i = 0; i < g_dbg_minv.size(); i++) minv[i] = g_dbg_minv[i];
  *nv = g_dbg_nv;
  return g_dbg_m;
}

int emu_solve_batch(const void* records, int batch, int h, int nv_cap, int m_cap, int max_iter, float* forces,
                    double* solution, int* info, double* H_out, double* g_out) {
  // MPC_EMU_PACKED=1: the packed-triangle layout of the nv <= 128 register class
  const bool packed = getenv("MPC_EMU_PACKED") && atoi(getenv("MPC_EMU_PACKED")) != 0;
  return packed ? emu_solve_batch_t<true>(records, batch, h, nv_cap, m_cap, max_iter, forces, solution, info, H_out, g_out)
                : emu_solve_batch_t<false>(records, batch, h, nv_cap, m_cap, max_iter, forces, solution, info, H_out, g_out);
}

// Wrench-space class (mpc_core.h): the same problems through H^{-1} = (I - G'MG)/(2 alpha), one emulated thread,
// full storage of the 6h x 6h matrix, generic sweep for the two inversions.
int emu_solve_batch_wrench(const void* records, int batch, int h, int m_cap, int max_iter, float* forces, double* solution,
                           int* info) {
  using namespace mpc;
  const int nv_cap = 12 * h;
  if (m_cap <= 0) m_cap = nv_cap;
  const Layout L = make_layout(h, nv_cap, m_cap, 1, 0, 0, 0, 1);
  std::vector<char> fast(L.fast_bytes + 64);
  const size_t stride = ((size_t)(4 * (MPC_REC_TRAJ + 12 * h) + 4 * h) + 15) / 16 * 16;
  OneThreadT<false, true> cx{0, 1};
  const int NU = 12 * h;
  for (int b = 0; b < batch; b++) {
    memset(fast.data(), 0xCD, fast.size());
    Work k = carve(L, fast.data(), nullptr);
    const float* rec = (const float*)((const char*)records + stride * b);
    const unsigned char* gait = (const unsigned char*)rec + 4 * (MPC_REC_TRAJ + 12 * h);
    k.i2a = 0.5 / (double)rec[MPC_REC_ALPHA];
    assemble_front(cx, rec, gait, k);
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      assemble_K(cx, rec, k);
      invert_spd(cx, k, 6 * h);
      if (k.sc->status == MPC_STATUS_OPTIMAL) {
        wr_form_second(cx, rec, k);
        invert_spd(cx, k, 6 * h);
      }
      if (k.sc->status == MPC_STATUS_OPTIMAL) {
        active_set_init(cx, rec, gait, k);
        active_set(cx, rec, gait, k, max_iter);
      }
    }
    int32_t st = 0;
    scatter(cx, k, forces + 12 * b, solution ? solution + (size_t)NU * b : nullptr, &st);
    if (info) {
      info[4 * b + 0] = k.sc->nv;
      info[4 * b + 1] = k.sc->m;
      info[4 * b + 2] = k.sc->iters;
      info[4 * b + 3] = k.sc->status;
    }
  }
  return 0;
}

// Riccati solver (csrc/mpc_riccati.h): the same problems without the condensed Hessian, one emulated thread.
// info [batch*4]: nv, active-set size at exit, iterations, status code (STATUS_RETRY_BIG = 0x40 when m_cap is hit).
int emu_solve_batch_riccati(const void* records, int batch, int h, int nv_cap, int m_cap, int max_iter, float* forces,
                            double* solution, int* info, int with_slab) {
  using namespace mpc;
  if (nv_cap <= 0) nv_cap = 12 * h;
  if (m_cap <= 0) m_cap = nv_cap;
  const RicLayout L = make_ric_layout(h, nv_cap, m_cap);
  std::vector<char> fast(L.bytes + 64);
  const size_t stride = ((size_t)(4 * (MPC_REC_TRAJ + 12 * h) + 4 * h) + 15) / 16 * 16;
  OneThreadT<false> cx{0, 1};
  const int NU = 12 * h;
  for (int b = 0; b < batch; b++) {
    memset(fast.data(), 0xCD, fast.size());  // poison: nothing may rely on zeroed workspace
    RicWork k = ric_carve(L, fast.data());
    std::vector<char> slab(with_slab ? L.slab_bytes : 0, (char)0xCD);
    if (with_slab) k.slab = slab.data();  // a working set that outgrows the tile moves here instead of being re-queued
    const float* rec = (const float*)((const char*)records + stride * b);
    const unsigned char* gait = (const unsigned char*)rec + 4 * (MPC_REC_TRAJ + 12 * h);
    ric_solve_problem(cx, rec, gait, k, max_iter);
    if (k.sc->status == MPC_STATUS_OPTIMAL && k.sc->nv > nv_cap) return -1;
    int32_t st = 0;
    ric_scatter(cx, k, forces + 12 * b, solution ? solution + (size_t)NU * b : nullptr, &st);
    if (info) {
      info[4 * b + 0] = k.sc->nv;
      info[4 * b + 1] = k.sc->m;
      info[4 * b + 2] = k.sc->iters;
      info[4 * b + 3] = k.sc->status;
    }
  }
  return 0;
}

// Warm start: warm_cache [batch][kWarmStride] ints, read before and rewritten after every solve.
int emu_solve_batch_warm(const void* records, int batch, int h, int nv_cap, int m_cap, int max_iter, float* forces,
                         double* solution, int* info, int* warm_cache, int warm_shift) {
  return emu_solve_batch_t<false>(records, batch, h, nv_cap, m_cap, max_iter, forces, solution, info, nullptr, nullptr,
                                  warm_cache, warm_shift);
}
int emu_warm_stride(void) { return mpc::kWarmStride; }

// Host build of the device-side record builder (csrc/mpc_ticks.h), one robot at a time.
void emu_build_records(const float* ticks, int batch, int h, unsigned char* records, float* state_out) {
  const size_t stride = ((size_t)(4 * (MPC_REC_TRAJ + 12 * h) + 4 * h) + 15) / 16 * 16;
  for (int b = 0; b < batch; b++)
    mpc::build_record_from_tick(ticks + (size_t)b * MPC_TICK_WORDS, h, (char*)records + stride * b, stride,
                                state_out ? state_out + 4 * b : nullptr);
}

// Layout check (no solve): byte ranges [begin, end) inside the fast workspace of every region a Work points to, for
// per-problem set `set` of make_layout(h, nv_cap, m_cap, 1, npad, packed, pipe).  out[2*i], out[2*i+1]; returns the
// number of regions, *fast_bytes the workspace size.  Order: sc, g, x, stance, posk, amask, W, Wia, Wiz, C, M, xs, qe,
// psum, mom, T, ck, ub, Wca, Wcz, w, r, u, tcol, red, Hm.
int emu_layout_regions(int h, int nv_cap, int m_cap, int npad, int packed, int pipe, int set, long* out, long* fast_bytes) {
  using namespace mpc;
  const Layout L = make_layout(h, nv_cap, m_cap, 1, npad, packed, pipe);
  std::vector<char> fast(L.fast_bytes + 64);
  char* base = fast.data();
  while (((uintptr_t)base & 15) != 0) base++;
  const Work k = carve(L, base, nullptr, set);
  *fast_bytes = L.fast_bytes;
  const int hm = packed ? nv_cap * (nv_cap + 1) / 2 : nv_cap * L.ld;
  struct R { const void* p; long bytes; };
  const R regs[] = {
      {k.sc, (long)sizeof(Scalars)}, {k.g, 8L * nv_cap}, {k.x, 8L * nv_cap}, {k.stance, 16L * h}, {k.posk, 16L * h},
      {k.amask, 16L * h}, {k.W, 4L * (m_cap + 1)}, {k.Wia, 4L * (m_cap + 1)}, {k.Wiz, 4L * (m_cap + 1)},
      {k.C, 8L * 3 * 156}, {k.M, 8L * 6 * 144}, {k.xs, 8L * 39}, {k.qe, 8L * 12 * h}, {k.psum, 8L * 5 * h},
      {k.mom, 8L * 3 * 12 * h}, {k.T, 8L * m_cap * L.ldT}, {k.ck, 8L * L.ck_len}, {k.ub, 8L * 4 * h},
      {k.Wca, 8L * (m_cap + 1)}, {k.Wcz, 8L * (m_cap + 1)}, {k.w, 8L * (m_cap + 1)}, {k.r, 8L * (m_cap + 1)},
      {k.u, 8L * (m_cap + 1)}, {k.tcol, 8L * (m_cap + 1)}, {k.red, 8L * kRedDoubles}, {k.Hm, 8L * hm}};
  const int n = (int)(sizeof(regs) / sizeof(regs[0]));
  for (int i = 0; i < n; i++) {
    out[2 * i] = (long)((const char*)regs[i].p - base);
    out[2 * i + 1] = out[2 * i] + regs[i].bytes;
  }
  return n;
}

}  // extern "C"

// Host build of the N2 / N4 device bodies (csrc/mpc_legs.h).
extern "C" void emu_gait_state(const int32_t* gait, int batch, float* state, unsigned char* tables, int table_stride) {
  for (int b = 0; b < batch; b++)
    mpc::gait_state_from_record(gait + (size_t)b * MPC_GAIT_WORDS, state + (size_t)b * MPC_GAIT_STATE_WORDS,
                                tables ? tables + (size_t)b * table_stride : nullptr);
}
extern "C" void emu_leg_commands(const float* legs, const float* forces, int batch, float* f_ff, float* tau) {
  for (int b = 0; b < batch; b++)
    mpc::leg_commands_from_record(legs + (size_t)b * MPC_LEG_WORDS, forces + (size_t)12 * b, f_ff + (size_t)12 * b,
                                  tau + (size_t)12 * b);
}
