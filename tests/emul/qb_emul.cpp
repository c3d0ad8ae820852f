// TEST INFRASTRUCTURE: a host emulator of k_fused_pass.
//
// Executes the planner's DevPass programs thread by thread exactly as the CUDA kernel is
// specified to (tile deposit, register/thread/external bit split, XOR-swizzled shared-memory
// transposes, gate predicates), so that planner and program-encoding bugs are caught on a
// CPU-only box, and so that the shared-memory bank behaviour of every transpose can be
// counted.  It is NOT part of libqubism_sv.so and nothing in qubism_b200 can reach it.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>

#include "../../qubism_b200/csrc/qb_internal.h"
#include "../../qubism_b200/csrc/qb_jit.h"

using namespace qb;

namespace {

struct EmuStats {
  int64_t passes = 0, rounds = 0, gates = 0, max_bank_conflict = 1, local_transposes = 0, transposes = 0;
};

void apply_gate_host(std::vector<double> &re, std::vector<double> &im, int R, const DevGate &g, uint32_t tid,
                     uint64_t basefull) {
  const bool ok_thr = ((tid & g.cthr) == g.cthr) && ((basefull & g.cext) == g.cext);
  const int NR = 1 << R;
  if (g.type == G_DIAG) {
    const bool sel_thr = ((tid & g.dthr) != 0) || ((basefull & g.dext) != 0);
    for (int i = 0; i < NR; ++i) {
      const bool one = sel_thr || ((uint32_t(i) & g.dreg) != 0);
      const double dr = one ? g.m[6] : g.m[0], di = one ? g.m[7] : g.m[1];
      if (ok_thr && ((uint32_t(i) & g.creg) == g.creg)) {
        const double xr = re[i], xi = im[i];
        re[i] = dr * xr - di * xi;
        im[i] = dr * xi + di * xr;
      }
    }
    return;
  }
  const int J = (int)(g.treg & 0xffu);
  if (J < 0 || J >= R) {
    std::fprintf(stderr, "emulator: gate target is not a register bit (treg=%u)\n", g.treg);
    re[0] = NAN;
    return;
  }
  for (int p = 0; p < NR / 2; ++p) {
    const int i0 = ((p >> J) << (J + 1)) | (p & ((1 << J) - 1));
    const int i1 = i0 | (1 << J);
    if (!(ok_thr && ((uint32_t(i0) & g.creg) == g.creg))) continue;
    const double x0r = re[i0], x0i = im[i0], x1r = re[i1], x1i = im[i1];
    if (g.type == G_SWAP) {
      re[i0] = x1r; im[i0] = x1i; re[i1] = x0r; im[i1] = x0i;
    } else if (g.type == G_ROT) {  // three shears with the kernel's coefficients (t, s)
      const double t = g.m[0], s = g.m[1];
      double a0r = std::fma(t, x1r, x0r), a0i = std::fma(t, x1i, x0i);
      const double a1r = std::fma(s, a0r, x1r), a1i = std::fma(s, a0i, x1i);
      a0r = std::fma(t, a1r, a0r);
      a0i = std::fma(t, a1i, a0i);
      re[i0] = a0r; im[i0] = a0i; re[i1] = a1r; im[i1] = a1i;
    } else if (g.type == G_REAL) {
      re[i0] = g.m[0] * x0r + g.m[2] * x1r;
      im[i0] = g.m[0] * x0i + g.m[2] * x1i;
      re[i1] = g.m[4] * x0r + g.m[6] * x1r;
      im[i1] = g.m[4] * x0i + g.m[6] * x1i;
    } else {
      re[i0] = g.m[0] * x0r - g.m[1] * x0i + g.m[2] * x1r - g.m[3] * x1i;
      im[i0] = g.m[0] * x0i + g.m[1] * x0r + g.m[2] * x1i + g.m[3] * x1r;
      re[i1] = g.m[4] * x0r - g.m[5] * x0i + g.m[6] * x1r - g.m[7] * x1i;
      im[i1] = g.m[4] * x0i + g.m[5] * x0r + g.m[6] * x1i + g.m[7] * x1r;
    }
  }
}

// One DevStep on one thread's registers, flip-mask semantics as in the kernel: register i
// holds logical register index i ^ f.
void apply_step_host(std::vector<double> &re, std::vector<double> &im, int R, const DevStep &S, uint32_t tid,
                     uint64_t basefull, uint32_t &f) {
  const int NR = 1 << R;
  for (int J = 0; J < R; ++J) {
    const uint32_t kind = (S.kinds >> (4 * J)) & 15u;
    if ((kind & SLOT_CLASS) == SLOT_NONE) continue;
    const bool fl = (f >> J) & 1u;
    if (fl && !(kind & SLOT_FLIP)) std::fprintf(stderr, "emulator: flip pending on a slot the planner marked flip-free\n");
    const double *m = S.slot[J];
    for (int p = 0; p < NR / 2; ++p) {
      const int i0 = ((p >> J) << (J + 1)) | (p & ((1 << J) - 1)), i1 = i0 | (1 << J);
      if ((kind & SLOT_CLASS) == SLOT_ROT) {
        const double t = fl ? -m[0] : m[0], s = fl ? -m[1] : m[1];
        re[i0] = std::fma(t, re[i1], re[i0]);
        im[i0] = std::fma(t, im[i1], im[i0]);
        re[i1] = std::fma(s, re[i0], re[i1]);
        im[i1] = std::fma(s, im[i0], im[i1]);
        re[i0] = std::fma(t, re[i1], re[i0]);
        im[i0] = std::fma(t, im[i1], im[i0]);
        continue;
      }
      // logical pair: register i holds logical index i ^ f
      const int l0 = fl ? i1 : i0, l1 = fl ? i0 : i1;
      const double x0r = re[l0], x0i = im[l0], x1r = re[l1], x1i = im[l1];
      if ((kind & SLOT_CLASS) == SLOT_GENERAL1 && (fl || m[0] != 1.0 || m[1] != 0.0))
        std::fprintf(stderr, "emulator: scaled-general slot with a pending flip or m00 != 1\n");
      if ((kind & SLOT_CLASS) == SLOT_REAL) {
        re[l0] = m[0] * x0r + m[1] * x1r;
        im[l0] = m[0] * x0i + m[1] * x1i;
        re[l1] = m[2] * x0r + m[3] * x1r;
        im[l1] = m[2] * x0i + m[3] * x1i;
      } else {
        re[l0] = m[0] * x0r - m[1] * x0i + m[2] * x1r - m[3] * x1i;
        im[l0] = m[0] * x0i + m[1] * x0r + m[2] * x1i + m[3] * x1r;
        re[l1] = m[4] * x0r - m[5] * x0i + m[6] * x1r - m[7] * x1i;
        im[l1] = m[4] * x0i + m[5] * x0r + m[6] * x1i + m[7] * x1r;
      }
    }
  }
  for (uint32_t k = 0; k < S.ntog; ++k) {
    const auto &tg = S.tog[k];
    if (((tid & tg.cthr) == tg.cthr) && ((basefull & tg.cext) == tg.cext)) f ^= 1u << tg.bit;
  }
  if (S.swap_j != 0xffu) {
    const int J = (int)(S.swap_j & 7u);
    if ((S.swap_j & 16u) && ((f & S.swap_creg) || S.swap_cthr || S.swap_cext || __builtin_popcount(S.swap_creg) != 1))
      std::fprintf(stderr, "emulator: static register swap flagged wrongly\n");
    const bool ok_thr = ((tid & S.swap_cthr) == S.swap_cthr) && ((basefull & S.swap_cext) == S.swap_cext);
    for (int p = 0; p < NR / 2; ++p) {
      const int i0 = ((p >> J) << (J + 1)) | (p & ((1 << J) - 1)), i1 = i0 | (1 << J);
      if (ok_thr && (((uint32_t(i0) ^ f) & S.swap_creg) == S.swap_creg)) {
        std::swap(re[i0], re[i1]);
        std::swap(im[i0], im[i1]);
      }
    }
  }
}

// end of round: put every amplitude into the register its logical index names
void fold_flip_host(std::vector<double> &re, std::vector<double> &im, int R, uint32_t f) {
  if (!f) return;
  const int NR = 1 << R;
  std::vector<double> r2(NR), i2(NR);
  for (int i = 0; i < NR; ++i) {
    r2[i ^ f] = re[i];
    i2[i ^ f] = im[i];
  }
  re = r2;
  im = i2;
}

uint32_t thread_u(const DevRound &rd, int nthr_bits, uint32_t tid) {
  uint32_t u = 0;
  for (int j = 0; j < nthr_bits; ++j) u |= ((tid >> j) & 1u) << rd.tid_pos[j];
  return u;
}

uint32_t reg_sx_of(const DevRound &rd, int R, int i) {
  uint32_t x = 0;
  for (int j = 0; j < R; ++j)
    if ((i >> j) & 1) x ^= rd.reg_sx[j];
  return x;
}

// worst bank-group multiplicity of one 128-bit shared access by a warp (quarter-warp phases)
int conflict_degree(const DevRound &rd, int T, int R, int i) {
  const int NT = 1 << (T - R);
  int worst = 1;
  for (int w = 0; w < NT; w += 8) {
    int cnt[8] = {0};
    for (int l = 0; l < 8 && w + l < NT; ++l) {
      const uint32_t x = swz_host(thread_u(rd, T - R, w + l)) ^ reg_sx_of(rd, R, i);
      worst = std::max(worst, ++cnt[x & 7]);
    }
  }
  return worst;
}

// xdst != nullptr and P.xch.n != 0: the stores carry a global<->local swap (XchGeom) -- tile by tile
// they go to (*xdst)[destination rank] exactly as the kernels address their peers' second shards;
// `amps` is left alone and the caller moves the buffers between the ranks.
void run_pass(const PassPlan &pp, int L, std::vector<double> &amps, EmuStats &st,
              std::vector<std::vector<double>> *xdst = nullptr) {
  const DevPass &P = *reinterpret_cast<const DevPass *>(pp.blob.data());
  const DevGate *G = pp.gates.data();
  const DevStep *S = reinterpret_cast<const DevStep *>(pp.blob.data() + sizeof(DevPass));
  const int T = P.tile_bits, R = P.reg_bits, NT = 1 << (T - R), NR = 1 << R;
  std::vector<double> tile_re(size_t(1) << T), tile_im(size_t(1) << T);
  std::vector<char> written(size_t(1) << T);
  std::vector<std::vector<double>> re(NT, std::vector<double>(NR)), im(NT, std::vector<double>(NR));
  for (uint32_t r = 1; r < P.nrounds; ++r)
    for (int i = 0; i < NR; ++i) {
      st.max_bank_conflict = std::max<int64_t>(st.max_bank_conflict, conflict_degree(P.rounds[r - 1], T, R, i));
      st.max_bank_conflict = std::max<int64_t>(st.max_bank_conflict, conflict_degree(P.rounds[r], T, R, i));
    }
  auto goff = [&](const DevRound &rd, uint32_t tid, int i, const uint8_t *pos) {
    uint64_t o = 0;
    for (int j = 0; j < T - R; ++j) o |= uint64_t((tid >> j) & 1u) << pos[rd.tid_pos[j]];
    for (int j = 0; j < R; ++j)
      if ((i >> j) & 1) o += 1ull << pos[rd.reg_pos[j]];
    return o;
  };
  // out of place: tile number t becomes block t of a second array, its bits permuted (out_pos)
  std::vector<double> other;
  if (P.oop) other.assign(amps.size(), NAN);
  std::vector<double> &dst = P.oop ? other : amps;
  for (uint64_t tile_id = 0; tile_id < pp.ntiles; ++tile_id) {
    uint64_t base = 0, t = tile_id;
    for (uint32_t k = 0; k < P.nruns; ++k) {
      base |= (t & ((1ull << P.run_len[k]) - 1)) << P.run_shift[k];
      t >>= P.run_len[k];
    }
    base |= P.base_fixed;
    uint64_t obase = base;
    if (P.oop) {
      obase = 0;
      uint64_t tt = tile_id;
      for (uint32_t k = 0; k < P.onruns; ++k) {
        obase |= (tt & ((1ull << P.orun_len[k]) - 1)) << P.orun_shift[k];
        tt >>= P.orun_len[k];
      }
    }
    const uint64_t basefull = base | P.rank_bits;
    for (int tid = 0; tid < NT; ++tid)
      for (int i = 0; i < NR; ++i) {
        const uint64_t a = base + goff(P.rounds[0], tid, i, P.tile_pos);
        re[tid][i] = amps[2 * a];
        im[tid][i] = amps[2 * a + 1];
      }
    for (uint32_t r = 0; r < P.nrounds; ++r) {
      const DevRound &RD = P.rounds[r];
      if (r > 0) {
        const DevRound &PR = P.rounds[r - 1];
        std::fill(written.begin(), written.end(), 0);
        // A warp-local transpose is emulated warp by warp against POISONED shared memory: if
        // the planner flagged it wrongly, a warp reads a slot another warp owns and gets NaN.
        const int group = RD.warp_local ? 32 : NT;
        if (RD.warp_local) {
          std::fill(tile_re.begin(), tile_re.end(), NAN);
          std::fill(tile_im.begin(), tile_im.end(), NAN);
        }
        for (int g0 = 0; g0 < NT; g0 += group) {
          for (int tid = g0; tid < g0 + group && tid < NT; ++tid)
            for (int i = 0; i < NR; ++i) {
              const uint32_t x = swz_host(thread_u(PR, T - R, tid)) ^ reg_sx_of(PR, R, i);
              if (written[x]) std::fprintf(stderr, "emulator: shared-memory slot written twice\n");
              written[x] = 1;
              tile_re[x] = re[tid][i];
              tile_im[x] = im[tid][i];
            }
          for (int tid = g0; tid < g0 + group && tid < NT; ++tid)
            for (int i = 0; i < NR; ++i) {
              const uint32_t x = swz_host(thread_u(RD, T - R, tid)) ^ reg_sx_of(RD, R, i);
              re[tid][i] = tile_re[x];
              im[tid][i] = tile_im[x];
            }
        }
      }
      if (P.lite) {
        // lite passes carry STEPS, not gates: run them exactly as the kernel does (rotation
        // slots in register-bit order, then the toggles, then the register-controlled X), with a
        // per-thread flip mask that folds into the data at the end of the round
        for (int tid = 0; tid < NT; ++tid) {
          uint32_t f = 0;
          for (uint32_t si = RD.step_begin; si < RD.step_end; ++si) apply_step_host(re[tid], im[tid], R, S[si], tid, basefull, f);
          fold_flip_host(re[tid], im[tid], R, f);
        }
      } else {
        for (uint32_t gi = RD.gate_begin; gi < RD.gate_end; ++gi)
          for (int tid = 0; tid < NT; ++tid) apply_gate_host(re[tid], im[tid], R, G[gi], tid, basefull);
      }
      st.gates += RD.gate_end - RD.gate_begin;
    }
    for (int tid = 0; tid < NT; ++tid)
      for (int i = 0; i < NR; ++i) {
        double xr = re[tid][i], xi = im[tid][i];
        if (P.has_gscale) {
          const double yr = P.gscale[0] * xr - P.gscale[1] * xi;
          xi = P.gscale[0] * xi + P.gscale[1] * xr;
          xr = yr;
        }
        uint64_t a = obase + goff(P.rounds[P.nrounds - 1], tid, i, P.out_pos);
        std::vector<double> *dstp = &dst;
        if (P.oop && P.xch.n) {  // (as k_fused_pass / the generated kernels compute it, store by store)
          uint32_t rr = P.xch.rbase;
          for (uint32_t k = 0; k < P.xch.n; ++k) rr |= uint32_t((a >> P.xch.lbit[k]) & 1ull) << P.xch.rbit[k];
          a = (a & ~P.xch.vmask) | P.xch.vconst;
          dstp = &(*xdst)[rr];
        }
        (*dstp)[2 * a] = xr;
        (*dstp)[2 * a + 1] = xi;
      }
  }
  if (P.oop && !P.xch.n) amps.swap(other);
  st.passes++;
  st.rounds += P.nrounds;
  for (uint32_t r = 1; r < P.nrounds; ++r) {
    st.transposes++;
    st.local_transposes += P.rounds[r].warp_local ? 1 : 0;
  }
  (void)L;
}

// Out-of-place passes leave the amplitudes in a new qubit layout (PlanResult::final_pos): put them
// back in index order, as the library's layout-aware read does.
void undo_layout(const PlanResult &plan, int nlocal, std::vector<double> &a) {
  if (plan.final_pos.empty()) return;
  std::vector<double> b(a.size());
  for (uint64_t x = 0; x < (1ull << nlocal); ++x) {
    uint64_t phys = 0;
    for (int bit = 0; bit < nlocal; ++bit)
      if ((x >> bit) & 1) phys |= 1ull << plan.final_pos[bit];
    b[2 * x] = a[2 * phys];
    b[2 * x + 1] = a[2 * phys + 1];
  }
  a.swap(b);
}

}  // namespace

// structural key -> host function of the generated code (g++ on the host flavour of the source),
// PROCESS-WIDE: if two different structures ever shared a key, a later circuit would run the wrong
// code here and fail its parity check -- the tests thereby also check that the key captures
// everything that shapes the generated source
typedef int (*host_fn)(double *, const double *, uint64_t, const void *, uint64_t);
static std::vector<std::pair<std::string, void *>> g_host_libs;
int host_code_for(const PassPlan &p, const JitProgram &kp, const char *workdir, host_fn *out) {
  static int nlibs = 0;  // (file names must never repeat inside a process: dlopen caches by path)
  void *fn = nullptr;
  for (auto &pr : g_host_libs)
    if (pr.first == kp.key) fn = pr.second;
  if (!fn) {
    JitProgram hp;
    std::string why;
    if (!jit_generate(p, JIT_HOST_SRC, hp, &why)) return -6;
    const std::string base = std::string(workdir) + "/qbj_" + std::to_string((long)getpid()) + "_" + std::to_string(nlibs++);
    FILE *f = std::fopen((base + ".cpp").c_str(), "w");
    if (!f) return -9;
    std::fwrite(hp.src.data(), 1, hp.src.size(), f);
    std::fclose(f);
    const std::string cmd = "g++ -std=c++17 -O1 -fPIC -shared -Wno-unknown-pragmas -o " + base + ".so " + base + ".cpp 2> " + base + ".log";
    if (std::system(cmd.c_str()) != 0) return -10;
    void *lib = dlopen((base + ".so").c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) return -11;
    fn = dlsym(lib, "qb_jit_pass_host");
    if (!fn) return -12;
    g_host_libs.emplace_back(kp.key, fn);
  }
  *out = reinterpret_cast<host_fn>(fn);
  return 0;
}

extern "C" {

// Queue `ops` (reference qubit indices) for an nlocal-qubit single-rank state, plan, and
// emulate the passes on `amps` (2 * 2^nlocal doubles, in place).  stats_out[0..3] = passes,
// rounds, gates executed, worst shared-memory bank multiplicity.  Returns 0, or -1 on error.
int qbe_run(int nlocal, const qb_op *ops, int64_t nops, const char *options, double *amps, int64_t *stats_out) {
  PlanOptions opt;
  if (options) {
    std::string o(options);
    size_t pos = 0;
    while (pos < o.size()) {
      size_t e = o.find(',', pos);
      if (e == std::string::npos) e = o.size();
      std::string kv = o.substr(pos, e - pos);
      size_t eq = kv.find('=');
      if (eq != std::string::npos && !set_opt(opt, kv.substr(0, eq), strtoll(kv.c_str() + eq + 1, nullptr, 10)))
        return -1;
      pos = e + 1;
    }
  }
  OpQueue q;
  q.reset(nlocal, opt.peephole != 0, opt.rot != 0);
  static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
  for (int64_t i = 0; i < nops; ++i) {
    const qb_op &o = ops[i];
    uint64_t cm = 0;
    const int nc = o.kind == 1 ? 1 : o.nctrl;
    for (int k = 0; k < nc; ++k) cm |= 1ull << (nlocal - 1 - o.ctrl[k]);
    q.push_1q(nlocal - 1 - o.target, cm, o.kind == 1 ? X : reinterpret_cast<const double *>(o.m));
  }
  int T, R;
  effective_tile(opt, nlocal, T, R);
  if (T == 0) return -2;  // the unfused kernels have no emulator (they are one-liners)
  opt.tile_bits = T;
  opt.reg_bits = R;
  if (!opt.fuse) opt.max_pass_gates = 1;
  if (opt.oop && opt.low_bits < opt.oop_low_bits) opt.low_bits = std::min(opt.oop_low_bits, T);
  std::vector<PhysOp> pops;
  for (const auto &h : q.ops) {
    if (h.dead) continue;
    PhysOp p;
    p.type = h.type;
    p.target = h.target;
    p.ctrl = h.ctrl;
    std::memcpy(p.m, h.m, sizeof(h.m));
    pops.push_back(p);
  }
  PlanResult plan = plan_passes(pops, nlocal, 0, opt, q.gscale);
  if (plan.consumed != pops.size()) return -3;
  std::vector<double> a(amps, amps + (size_t(2) << nlocal));
  EmuStats st;
  for (const auto &p : plan.passes) run_pass(p, nlocal, a, st);
  undo_layout(plan, nlocal, a);
  if (plan.passes.empty() && !(q.gscale[0] == 1.0 && q.gscale[1] == 0.0)) {
    for (size_t i = 0; i < (size_t(1) << nlocal); ++i) {
      const double xr = a[2 * i], xi = a[2 * i + 1];
      a[2 * i] = q.gscale[0] * xr - q.gscale[1] * xi;
      a[2 * i + 1] = q.gscale[0] * xi + q.gscale[1] * xr;
    }
  }
  std::memcpy(amps, a.data(), sizeof(double) * a.size());
  if (stats_out) {
    stats_out[0] = st.passes;
    stats_out[1] = st.rounds;
    stats_out[2] = st.gates;
    stats_out[3] = st.max_bank_conflict;
    stats_out[4] = st.transposes;
    stats_out[5] = st.local_transposes;
  }
  return 0;
}

// The structure-specialised kernels (qb_jit.cpp) on the CPU: plan as qbe_run does, then run every
// pass the generator accepts through the HOST emulation of its generated source (compiled with
// g++ into `workdir`, loaded with dlopen), every other pass through the emulator above.  The
// factors the specialised passes leave out are applied at the end, as the state's deferred
// scalar would be.  stats_out[0] = passes, [1] = specialised passes, [2] = distinct structures.
// If dev_src_out is not null, the DEVICE source of specialised pass number `dev_src_index` is
// copied there (for an NVRTC compile check that needs no GPU).
int qbe_run_jit(int nlocal, const qb_op *ops, int64_t nops, const char *options, double *amps, int64_t *stats_out,
                const char *workdir, int dev_src_index, char *dev_src_out, int64_t dev_src_cap) {
  PlanOptions opt;
  if (options) {
    std::string o(options);
    size_t pos = 0;
    while (pos < o.size()) {
      size_t e = o.find(',', pos);
      if (e == std::string::npos) e = o.size();
      std::string kv = o.substr(pos, e - pos);
      size_t eq = kv.find('=');
      if (eq != std::string::npos && !set_opt(opt, kv.substr(0, eq), strtoll(kv.c_str() + eq + 1, nullptr, 10)))
        return -1;
      pos = e + 1;
    }
  }
  OpQueue q;
  q.reset(nlocal, opt.peephole != 0, opt.rot != 0);
  static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
  for (int64_t i = 0; i < nops; ++i) {
    const qb_op &o = ops[i];
    uint64_t cm = 0;
    const int nc = o.kind == 1 ? 1 : o.nctrl;
    for (int k = 0; k < nc; ++k) cm |= 1ull << (nlocal - 1 - o.ctrl[k]);
    q.push_1q(nlocal - 1 - o.target, cm, o.kind == 1 ? X : reinterpret_cast<const double *>(o.m));
  }
  int T, R;
  effective_tile(opt, nlocal, T, R);
  if (T == 0) return -2;
  opt.tile_bits = T;
  opt.reg_bits = R;
  if (opt.oop && opt.low_bits < opt.oop_low_bits) opt.low_bits = std::min(opt.oop_low_bits, T);
  std::vector<PhysOp> pops;
  for (const auto &h : q.ops) {
    if (h.dead) continue;
    PhysOp p;
    p.type = h.type;
    p.target = h.target;
    p.ctrl = h.ctrl;
    std::memcpy(p.m, h.m, sizeof(h.m));
    pops.push_back(p);
  }
  PlanResult plan = plan_passes(pops, nlocal, 0, opt, q.gscale);
  if (plan.consumed != pops.size()) return -3;
  std::vector<double> a(amps, amps + (size_t(2) << nlocal));
  EmuStats st;
  double pending = 1.0;
  int njit = 0;
  auto &libs = g_host_libs;
  const size_t libs_before = libs.size();
  int nguard = 0;
  for (auto &p : plan.passes) {
    DevPass &P = *reinterpret_cast<DevPass *>(p.blob.data());
    {
      // the flush's range guard (qb_api.cpp run_fused_segment): when the factor the specialised
      // passes have left out so far would fall below 2^-300 with this pass, the pass applies the
      // running factor on its way out (has_gscale is structural: a different kernel)
      JitProgram probe;
      if (!P.has_gscale && jit_quick(p, probe, nullptr) && !(std::fabs(pending * probe.left_out) > 0x1p-300)) {
        P.has_gscale = 1;
        P.gscale[0] = pending * probe.left_out;
        P.gscale[1] = 0.0;
        pending = 1.0 / probe.left_out;  // (multiplied by left_out again below: 1 after this pass)
        ++nguard;
      }
    }
    JitProgram kp, full_key;
    std::string why;
    // the flush uses jit_quick (digest + coefficients, no strings); check it against the full walk
    const bool ok_quick = jit_quick(p, kp, &why);
    const bool ok_full = jit_generate(p, JIT_KEY_ONLY, full_key, &why);
    if (ok_quick != ok_full) return -20;
    if (ok_quick && (kp.coefs != full_key.coefs || kp.left_out != full_key.left_out || kp.args_bytes != full_key.args_bytes))
      return -21;
    if (!ok_quick) {
      run_pass(p, nlocal, a, st);
      continue;
    }
    if (dev_src_out && njit == dev_src_index) {
      JitProgram dp;
      if (!jit_generate(p, JIT_DEVICE_SRC, dp, &why)) return -6;
      if ((int64_t)dp.src.size() + 1 > dev_src_cap) return -7;
      std::memcpy(dev_src_out, dp.src.c_str(), dp.src.size() + 1);
      if (dp.key.size() == 0 || dp.coefs != kp.coefs) return -8;
    }
    host_fn fn = nullptr;
    if (int rc = host_code_for(p, kp, workdir, &fn)) return rc;
    const double one[2] = {1.0, 0.0};
    const std::vector<uint8_t> args = jit_pack_args(kp, P.has_gscale ? P.gscale : one, P.rank_bits, P.base_fixed);
    if (P.oop) {
      std::vector<double> other(a.size(), NAN);
      if (fn(other.data(), a.data(), p.ntiles, args.data(), (uint64_t)args.size()) != 0) return -13;
      a.swap(other);
    } else if (fn(a.data(), a.data(), p.ntiles, args.data(), (uint64_t)args.size()) != 0) {
      return -13;
    }
    pending *= kp.left_out;
    ++njit;
    st.passes++;
  }
  undo_layout(plan, nlocal, a);
  const bool qg = plan.passes.empty() && !(q.gscale[0] == 1.0 && q.gscale[1] == 0.0);
  for (size_t i = 0; i < (size_t(1) << nlocal); ++i) {
    double xr = a[2 * i] * pending, xi = a[2 * i + 1] * pending;
    if (qg) {
      const double yr = q.gscale[0] * xr - q.gscale[1] * xi;
      xi = q.gscale[0] * xi + q.gscale[1] * xr;
      xr = yr;
    }
    a[2 * i] = xr;
    a[2 * i + 1] = xi;
  }
  std::memcpy(amps, a.data(), sizeof(double) * a.size());
  if (stats_out) {
    stats_out[0] = st.passes;
    stats_out[1] = njit;
    stats_out[2] = (int64_t)(libs.size() - libs_before);
    stats_out[3] = nguard;
  }
  return 0;
}


// Plan only (any nlocal, no amplitudes): write the DEVICE source of every pass the generator
// accepts to `outdir`/pass_<i>.cu.  Returns the number of passes planned, or a negative error;
// stats_out[0] = specialised passes, [1] = rotations counted from the coefficient vectors.
int qbe_jit_dump(int nlocal, const qb_op *ops, int64_t nops, const char *options, const char *outdir, int64_t *stats_out) {
  PlanOptions opt;
  if (options) {
    std::string o(options);
    size_t pos = 0;
    while (pos < o.size()) {
      size_t e = o.find(',', pos);
      if (e == std::string::npos) e = o.size();
      std::string kv = o.substr(pos, e - pos);
      size_t eq = kv.find('=');
      if (eq != std::string::npos && !set_opt(opt, kv.substr(0, eq), strtoll(kv.c_str() + eq + 1, nullptr, 10)))
        return -1;
      pos = e + 1;
    }
  }
  OpQueue q;
  q.reset(nlocal, opt.peephole != 0, opt.rot != 0);
  static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
  for (int64_t i = 0; i < nops; ++i) {
    const qb_op &o = ops[i];
    uint64_t cm = 0;
    const int nc = o.kind == 1 ? 1 : o.nctrl;
    for (int k = 0; k < nc; ++k) cm |= 1ull << (nlocal - 1 - o.ctrl[k]);
    q.push_1q(nlocal - 1 - o.target, cm, o.kind == 1 ? X : reinterpret_cast<const double *>(o.m));
  }
  int T, R;
  effective_tile(opt, nlocal, T, R);
  if (T == 0) return -2;
  opt.tile_bits = T;
  opt.reg_bits = R;
  if (opt.oop && opt.low_bits < opt.oop_low_bits) opt.low_bits = std::min(opt.oop_low_bits, T);
  std::vector<PhysOp> pops;
  for (const auto &h : q.ops) {
    if (h.dead) continue;
    PhysOp p;
    p.type = h.type;
    p.target = h.target;
    p.ctrl = h.ctrl;
    std::memcpy(p.m, h.m, sizeof(h.m));
    pops.push_back(p);
  }
  PlanResult plan = plan_passes(pops, nlocal, 0, opt, q.gscale);
  if (plan.consumed != pops.size()) return -3;
  int njit = 0, idx = 0;
  int64_t nfixed = 0, nconf = 0, ntrans = 0;
  if (std::getenv("QBE_TIME_KEYS")) {  // host cost of the structural key (what every flush pays per pass)
    const auto t0 = std::chrono::steady_clock::now();
    size_t bytes = 0;
    for (int rep = 0; rep < 50; ++rep)
      for (const auto &p : plan.passes) {
        JitProgram kp;
        if (jit_generate(p, JIT_KEY_ONLY, kp, nullptr)) bytes += kp.key.size();
      }
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "key-only generation: %.1f us per pass, %.0f bytes per key\n", us / (50.0 * plan.passes.size()),
                 double(bytes) / (50.0 * plan.passes.size()));
    const auto t1 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < 50; ++rep)
      for (const auto &p : plan.passes) {
        JitProgram kp;
        if (jit_quick(p, kp, nullptr)) bytes += kp.key.size();
      }
    const double us2 = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t1).count();
    std::fprintf(stderr, "jit_quick: %.2f us per pass\n", us2 / (50.0 * plan.passes.size()));
  }
  for (const auto &p : plan.passes) {
    JitProgram dp;
    std::string why;
    if (jit_generate(p, JIT_DEVICE_SRC, dp, &why)) {
      const std::string path = std::string(outdir) + "/pass_" + std::to_string(idx) + ".cu";
      FILE *f = std::fopen(path.c_str(), "w");
      if (!f) return -9;
      std::fwrite(dp.src.data(), 1, dp.src.size(), f);
      std::fclose(f);
      ++njit;
      nfixed += dp.swz_fixed;
      nconf += dp.swz_conflicts;
      ntrans += dp.nrounds - 1;
    }
    ++idx;
  }
  if (stats_out) {
    stats_out[0] = njit;
    stats_out[1] = ntrans;
    stats_out[2] = nfixed;
    stats_out[3] = nconf;
  }
  return idx;
}


// Layout dynamics of a sharded state under a REPEATED op stream (host only, no amplitudes): plan
// and choose swaps exactly as a flush would, `nsteps` times in a row; out_perm receives the
// logical->physical map after every step (nsteps x n ints), out_counts per step
// (passes, swaps, structures not seen in any earlier step).
static int64_t g_trace_fused = 0;  // swaps of the last qbe_layout_trace call that a pass's stores could carry
int64_t qbe_trace_fused() { return g_trace_fused; }
int qbe_layout_trace(int n, int nranks, const qb_op *ops, int64_t nops, const char *options, int nsteps, int any_local,
                     int *out_perm, int64_t *out_counts) {
  PlanOptions opt;
  if (options) {
    std::string o(options);
    size_t pos = 0;
    while (pos < o.size()) {
      size_t e = o.find(',', pos);
      if (e == std::string::npos) e = o.size();
      std::string kv = o.substr(pos, e - pos);
      size_t eq = kv.find('=');
      if (eq != std::string::npos && !set_opt(opt, kv.substr(0, eq), strtoll(kv.c_str() + eq + 1, nullptr, 10)))
        return -1;
      pos = e + 1;
    }
  }
  int pbits = 0;
  while ((1 << pbits) < nranks) ++pbits;
  const int L = n - pbits;
  int T, R;
  effective_tile(opt, L, T, R);
  if (T == 0) return -2;
  opt.tile_bits = T;
  opt.reg_bits = R;
  if (nranks > 1 && !opt.oop_dist) opt.oop = 0;  // (as in qb_api.cpp)
  if (opt.oop && opt.low_bits < opt.oop_low_bits) opt.low_bits = std::min(opt.oop_low_bits, T);
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  std::vector<std::string> seen_keys;
  static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
  g_trace_fused = 0;
  for (int step = 0; step < nsteps; ++step) {
    OpQueue q;
    q.reset(n, opt.peephole != 0, opt.rot != 0);
    for (int64_t i = 0; i < nops; ++i) {
      const qb_op &o = ops[i];
      uint64_t cm = 0;
      const int nc = o.kind == 1 ? 1 : o.nctrl;
      for (int k = 0; k < nc; ++k) cm |= 1ull << (n - 1 - o.ctrl[k]);
      q.push_1q(n - 1 - o.target, cm, o.kind == 1 ? X : reinterpret_cast<const double *>(o.m));
    }
    std::vector<const HostOp *> seg;
    for (const auto &h : q.ops)
      if (!h.dead) seg.push_back(&h);
    opt.layout_known = 0;
    int64_t npass = 0, nswap = 0, nnew = 0;
    std::vector<std::string> step_keys;
    while (!seg.empty()) {
      std::vector<PhysOp> pops(seg.size());
      for (size_t i = 0; i < seg.size(); ++i) {
        const HostOp &h = *seg[i];
        pops[i].type = h.type;
        pops[i].target = perm[h.target];
        uint64_t cm = 0;
        for (uint64_t b = h.ctrl; b; b &= b - 1) cm |= 1ull << perm[__builtin_ctzll(b)];
        pops[i].ctrl = cm;
        std::memcpy(pops[i].m, h.m, sizeof(h.m));
      }
      std::vector<int> labels(L, 0);  // the logical bit on each local physical bit
      for (int qq = 0; qq < n; ++qq)
        if (perm[qq] < L) labels[perm[qq]] = qq;
      PlanResult plan = plan_passes_until_swap(pops, L, 0, opt, &labels);
      if (getenv("QBE_TRACE_PLAN")) {
        std::fprintf(stderr, "step %d: plan of %zu ops -> %zu passes, gates:", step, seg.size(), plan.passes.size());
        for (const auto &p : plan.passes) std::fprintf(stderr, " %u/%ur", (unsigned)p.ngates, (unsigned)p.nrounds);
        std::fprintf(stderr, " consumed %zu\n", plan.consumed);
      }
      for (const auto &p : plan.passes) {
        ++npass;
        JitProgram kp;
        if (any_local & 4) {
          // LOGICAL signature (experiment): rounds and gates with physical positions replaced by the
          // logical qubits sitting there -- does the plan recur up to a relabelling of positions?
          std::vector<int> logical_of(n, -1);
          for (int qq = 0; qq < n; ++qq) logical_of[perm[qq]] = qq;
          const DevPass &P = *reinterpret_cast<const DevPass *>(p.blob.data());
          std::string sig;
          for (uint32_t r = 0; r < P.nrounds; ++r) {
            sig += "R";
            for (int j = 0; j < (int)P.reg_bits; ++j) sig += std::to_string(logical_of[P.tile_pos[P.rounds[r].reg_pos[j]]]) + ",";
            sig += "T";
            for (int j = 0; j < (int)(P.tile_bits - P.reg_bits); ++j)
              sig += std::to_string(logical_of[P.tile_pos[P.rounds[r].tid_pos[j]]]) + ",";
          }
          for (int oi : p.op_index) sig += "g" + std::to_string(oi) + ",";
          step_keys.push_back(sig);
        } else if (jit_generate(p, JIT_KEY_ONLY, kp, nullptr)) {
          step_keys.push_back(kp.key);
        }
      }
      if (!plan.final_pos.empty()) {  // out-of-place passes moved the local qubits
        for (int &x : perm)
          if (x < L) x = plan.final_pos[x];
        opt.layout_known = 1;
      }
      if (plan.consumed == seg.size()) break;
      std::vector<const HostOp *> rest;
      for (size_t i = 0; i < seg.size(); ++i)
        if (!plan.done[i]) rest.push_back(seg[i]);
      std::vector<const HostOp *> all;
      for (const auto &h : q.ops)
        if (!h.dead) all.push_back(&h);
      std::vector<SwapPair> sw = choose_swaps(n, L, perm, rest, (any_local & 1) != 0, (any_local & 2) ? &all : nullptr);
      if (sw.empty()) return -4;
      if (!plan.passes.empty() && !plan.final_pos.empty()) {  // would the last pass's stores carry this swap? (option fuse_exchange)
        XchGeom X;
        const DevPass &LP = *reinterpret_cast<const DevPass *>(plan.passes.back().blob.data());
        if (fused_exchange_geometry(LP, L, 0, nranks, sw, &X)) ++g_trace_fused;
        else if (getenv("QBE_TRACE_WHY")) {
          std::fprintf(stderr, "step %d: not fusable: oop=%d T=%d pairs", step, (int)LP.oop, (int)LP.tile_bits);
          for (const SwapPair &sp : sw) std::fprintf(stderr, " (g%d,l%d)", sp.gbit, sp.lbit);
          std::fprintf(stderr, "\n");
        }
      } else if (getenv("QBE_TRACE_WHY")) {
        std::fprintf(stderr, "step %d: no pass before the swap (passes %zu, final_pos %zu)\n", step, plan.passes.size(), plan.final_pos.size());
      }
      apply_swaps_to_perm(perm, sw);
      ++nswap;
      seg.swap(rest);
    }
    for (const auto &k : step_keys)
      if (std::find(seen_keys.begin(), seen_keys.end(), k) == seen_keys.end()) ++nnew;
    for (auto &k : step_keys) seen_keys.push_back(std::move(k));
    for (int i = 0; i < n; ++i) out_perm[step * n + i] = perm[i];
    out_counts[step * 3 + 0] = npass;
    out_counts[step * 3 + 1] = nswap;
    out_counts[step * 3 + 2] = nnew;
  }
  return 0;
}


// Distributed flush of ONE rank, emulated on the host.  `amps` is this rank's shard (2 * 2^L
// doubles).  Whenever the planner is stuck on gates that target global qubits, the same
// choose_swaps / swap_schedule code the NCCL path uses decides the exchange, and `xchg` moves
// the data: xchg(peer, send, recv, ndoubles) must exchange the buffers with `peer`
// (tests implement it with torch.distributed send/recv over gloo).  perm_inout[n] is the
// logical->physical bit map.  Returns the number of swaps, or a negative error.
typedef int (*qbe_xchg_fn)(int peer, const double *send, double *recv, int64_t ndoubles);

int qbe_run_rank(int n, int nranks, int rank, const qb_op *ops, int64_t nops, const char *options, double *amps,
                 int *perm_inout, qbe_xchg_fn xchg, int64_t *stats_out, int any_local) {
  PlanOptions opt;
  if (options) {
    std::string o(options);
    size_t pos = 0;
    while (pos < o.size()) {
      size_t e = o.find(',', pos);
      if (e == std::string::npos) e = o.size();
      std::string kv = o.substr(pos, e - pos);
      size_t eq = kv.find('=');
      if (eq != std::string::npos && !set_opt(opt, kv.substr(0, eq), strtoll(kv.c_str() + eq + 1, nullptr, 10)))
        return -1;
      pos = e + 1;
    }
  }
  int pbits = 0;
  while ((1 << pbits) < nranks) ++pbits;
  const int L = n - pbits;
  OpQueue q;
  q.reset(n, opt.peephole != 0, opt.rot != 0);
  static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
  for (int64_t i = 0; i < nops; ++i) {
    const qb_op &o = ops[i];
    uint64_t cm = 0;
    const int nc = o.kind == 1 ? 1 : o.nctrl;
    for (int k = 0; k < nc; ++k) cm |= 1ull << (n - 1 - o.ctrl[k]);
    q.push_1q(n - 1 - o.target, cm, o.kind == 1 ? X : reinterpret_cast<const double *>(o.m));
  }
  int T, R;
  if (!opt.oop_dist) opt.oop = 0;  // (as in qb_api.cpp)
  effective_tile(opt, L, T, R);
  if (T == 0) return -2;
  opt.tile_bits = T;
  opt.reg_bits = R;
  if (opt.oop && opt.low_bits < opt.oop_low_bits) opt.low_bits = std::min(opt.oop_low_bits, T);
  std::vector<int> perm(perm_inout, perm_inout + n);
  std::vector<const HostOp *> seg;
  for (const auto &h : q.ops)
    if (!h.dead) seg.push_back(&h);
  std::vector<double> a(amps, amps + (size_t(2) << L));
  EmuStats st;
  int nswaps = 0;
  int64_t nfused = 0, njit_fused = 0;
  bool gdone = (q.gscale[0] == 1.0 && q.gscale[1] == 0.0);
  while (!seg.empty()) {
    std::vector<PhysOp> pops(seg.size());
    for (size_t i = 0; i < seg.size(); ++i) {
      const HostOp &h = *seg[i];
      pops[i].type = h.type;
      pops[i].target = perm[h.target];
      uint64_t cm = 0;
      for (uint64_t b = h.ctrl; b; b &= b - 1) cm |= 1ull << perm[__builtin_ctzll(b)];
      pops[i].ctrl = cm;
      std::memcpy(pops[i].m, h.m, sizeof(h.m));
    }
    std::vector<int> labels(L, 0);
    for (int qq = 0; qq < n; ++qq)
      if (perm[qq] < L) labels[perm[qq]] = qq;
    PlanResult plan = plan_passes_until_swap(pops, L, rank, opt, &labels);
    const bool all = plan.consumed == seg.size();
    if (all && !gdone && !plan.passes.empty()) {
      DevPass *P = reinterpret_cast<DevPass *>(plan.passes.back().blob.data());
      P->gscale[0] = q.gscale[0];
      P->gscale[1] = q.gscale[1];
      P->has_gscale = 1;
      gdone = true;
    }
    std::vector<const HostOp *> rest;
    for (size_t i = 0; i < seg.size(); ++i)
      if (!plan.done[i]) rest.push_back(seg[i]);
    // any_local: bit 0 = Belady over any local bit (peer-memory path), bit 1 = cyclic tie-break
    // (the flush's own op stream as the future, as qb_api.cpp's make_local passes it), bit 2 = the
    // swap rides on the stores of the plan's last pass where it can (option fuse_exchange)
    std::vector<const HostOp *> all_ops;
    for (const auto &h : q.ops)
      if (!h.dead) all_ops.push_back(&h);
    std::vector<SwapPair> sw;
    bool fused = false;
    if (!all) {
      std::vector<int> perm2 = perm;  // the layout after this plan: what the swap is chosen on
      if (!plan.final_pos.empty())
        for (int &x : perm2)
          if (x < L) x = plan.final_pos[x];
      sw = choose_swaps(n, L, perm2, rest, (any_local & 1) != 0, (any_local & 2) ? &all_ops : nullptr);
      if (sw.empty()) return -4;
      if ((any_local & 4) && !plan.passes.empty() && !plan.final_pos.empty()) {
        DevPass *LP = reinterpret_cast<DevPass *>(plan.passes.back().blob.data());
        fused = fused_exchange_geometry(*LP, L, rank, nranks, sw, &LP->xch);
      }
    }
    for (size_t pi = 0; pi < plan.passes.size(); ++pi) {
      if (!fused || pi + 1 < plan.passes.size()) {
        run_pass(plan.passes[pi], L, a, st);
        continue;
      }
      // the last pass stores into one buffer per destination rank; gloo then stands in for NVLink:
      // what I wrote for `peer` goes there, what `peer` wrote for me fills the places whose victim
      // bits carry the peer's old rank-bit values
      const DevPass &LP = *reinterpret_cast<const DevPass *>(plan.passes[pi].blob.data());
      std::vector<std::vector<double>> xdst(nranks);
      for (uint32_t sel = 0; sel < (1u << LP.xch.n); ++sel) {
        uint32_t rr = LP.xch.rbase;
        for (uint32_t j = 0; j < LP.xch.n; ++j) rr |= ((sel >> j) & 1u) << LP.xch.rbit[j];
        xdst[rr].assign(a.size(), NAN);
      }
      JitProgram kp;
      if ((any_local & 8) && jit_quick(plan.passes[pi], kp, nullptr)) {
        // the same pass as GENERATED code (host flavour; the peer table holds the per-rank buffers):
        // what it leaves out is applied right away
        DevPass &XP = *reinterpret_cast<DevPass *>(plan.passes[pi].blob.data());
        for (int r = 0; r < nranks; ++r) XP.xch.peer[r] = reinterpret_cast<uint64_t>(xdst[r].data());
        host_fn fn = nullptr;
        const char *wd = getenv("QBE_WORKDIR");
        if (int rc = host_code_for(plan.passes[pi], kp, wd ? wd : "/tmp", &fn)) return rc;
        const double one[2] = {1.0, 0.0};
        const std::vector<uint8_t> args = jit_pack_args(kp, XP.has_gscale ? XP.gscale : one, XP.rank_bits, XP.base_fixed, &XP.xch);
        if (fn(nullptr, a.data(), plan.passes[pi].ntiles, args.data(), (uint64_t)args.size()) != 0) return -13;
        for (auto &v : xdst)
          for (double &x : v) x *= kp.left_out;  // (NaN stays NaN)
        st.passes++;
        ++njit_fused;
      } else {
        run_pass(plan.passes[pi], L, a, st, &xdst);
      }
      const int k = (int)sw.size();
      const uint64_t block = 1ull << (L - k);
      auto deposit = [&](uint64_t t) {
        uint64_t idx = 0;
        int src = 0;
        for (int b2 = 0; b2 < L; ++b2)
          if (!(LP.xch.vmask & (1ull << b2))) {
            idx |= ((t >> src) & 1ull) << b2;
            ++src;
          }
        return idx;
      };
      auto vbits_of = [&](int r) {  // rank r's values on the swapped rank bits, placed on the victims' positions
        uint64_t v = 0;
        for (uint32_t j = 0; j < LP.xch.n; ++j) v |= uint64_t((r >> LP.xch.rbit[j]) & 1) << LP.xch.lbit[j];
        return v;
      };
      std::vector<double> b(a.size(), NAN), send(2 * block), recv(2 * block);
      for (uint32_t sel = 0; sel < (1u << LP.xch.n); ++sel) {
        int peer = rank;
        for (uint32_t j = 0; j < LP.xch.n; ++j) peer ^= int((sel >> j) & 1u) << LP.xch.rbit[j];
        for (uint64_t t = 0; t < block; ++t) {
          const uint64_t i = deposit(t) | LP.xch.vconst;
          send[2 * t] = xdst[peer][2 * i];
          send[2 * t + 1] = xdst[peer][2 * i + 1];
        }
        if (peer == rank) recv = send;
        else if (xchg(peer, send.data(), recv.data(), (int64_t)(2 * block)) != 0) return -5;
        const uint64_t at = vbits_of(peer);
        for (uint64_t t = 0; t < block; ++t) {
          const uint64_t i = deposit(t) | at;
          b[2 * i] = recv[2 * t];
          b[2 * i + 1] = recv[2 * t + 1];
        }
      }
      // nothing else may have been written: every other place of the per-rank buffers is still NaN
      for (int r = 0; r < nranks; ++r)
        for (size_t i = 0; i < xdst[r].size() / 2; ++i)
          if ((i & LP.xch.vmask) != LP.xch.vconst && !std::isnan(xdst[r][2 * i])) return -6;
      a.swap(b);
      ++nfused;
    }
    if (!plan.final_pos.empty()) {
      for (int &x : perm)
        if (x < L) x = plan.final_pos[x];
      opt.layout_known = 1;
    }
    if (all) break;
    if (fused) {
      apply_swaps_to_perm(perm, sw);
      ++nswaps;
      seg.swap(rest);
      continue;
    }
    const int k = (int)sw.size();
    const uint64_t block = 1ull << (L - k);
    uint64_t swapped = 0;
    for (const SwapPair &sp : sw) swapped |= 1ull << sp.lbit;
    std::vector<double> send(2 * block), recv(2 * block);
    auto deposit = [&](uint64_t t) {  // free index -> local index with the swapped bits clear
      uint64_t idx = 0;
      int src = 0;
      for (int b2 = 0; b2 < L; ++b2)
        if (!(swapped & (1ull << b2))) {
          idx |= ((t >> src) & 1ull) << b2;
          ++src;
        }
      return idx;
    };
    for (const SwapStep &stp : swap_schedule(rank, L, sw)) {
      // my elements whose swapped local bits equal the peer's rank-bit value leave; the peer's
      // elements whose swapped bits equal MINE arrive and take exactly those places
      const uint64_t mine_at = place_sel(stp.my_sel, sw);
      for (uint64_t t = 0; t < block; ++t) {
        const uint64_t i = deposit(t) | mine_at;
        send[2 * t] = a[2 * i];
        send[2 * t + 1] = a[2 * i + 1];
      }
      if (xchg(stp.peer, send.data(), recv.data(), (int64_t)(2 * block)) != 0) return -5;
      for (uint64_t t = 0; t < block; ++t) {
        const uint64_t i = deposit(t) | mine_at;
        a[2 * i] = recv[2 * t];
        a[2 * i + 1] = recv[2 * t + 1];
      }
    }
    apply_swaps_to_perm(perm, sw);
    ++nswaps;
    seg.swap(rest);
  }
  if (!gdone)
    for (size_t i = 0; i < (size_t(1) << L); ++i) {
      const double xr = a[2 * i], xi = a[2 * i + 1];
      a[2 * i] = q.gscale[0] * xr - q.gscale[1] * xi;
      a[2 * i + 1] = q.gscale[0] * xi + q.gscale[1] * xr;
    }
  std::memcpy(amps, a.data(), sizeof(double) * a.size());
  for (int i = 0; i < n; ++i) perm_inout[i] = perm[i];
  if (stats_out) {
    stats_out[0] = st.passes;
    stats_out[1] = st.rounds;
    stats_out[2] = st.gates;
    stats_out[3] = st.max_bank_conflict;
    stats_out[4] = nfused;
    stats_out[5] = njit_fused;
  }
  return nswaps;
}

// Device source of a pass whose stores carry a global<->local swap (option fuse_exchange), for an
// NVRTC compile check that needs no GPU: plan `ops` on an nlocal-qubit shard of rank 1 of 4, hand the
// last out-of-place pass a swap of the two rank bits with local bits `lbit0`, `lbit1`, generate.
int qbe_xch_source(int nlocal, const qb_op *ops, int64_t nops, const char *options, int lbit0, int lbit1, char *out, int64_t cap) {
  PlanOptions opt;
  if (options) {
    std::string o(options);
    size_t pos = 0;
    while (pos < o.size()) {
      size_t e = o.find(',', pos);
      if (e == std::string::npos) e = o.size();
      std::string kv = o.substr(pos, e - pos);
      size_t eq = kv.find('=');
      if (eq != std::string::npos && !set_opt(opt, kv.substr(0, eq), strtoll(kv.c_str() + eq + 1, nullptr, 10)))
        return -1;
      pos = e + 1;
    }
  }
  OpQueue q;
  q.reset(nlocal, opt.peephole != 0, opt.rot != 0);
  static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
  for (int64_t i = 0; i < nops; ++i) {
    const qb_op &o = ops[i];
    uint64_t cm = 0;
    const int nc = o.kind == 1 ? 1 : o.nctrl;
    for (int k = 0; k < nc; ++k) cm |= 1ull << (nlocal - 1 - o.ctrl[k]);
    q.push_1q(nlocal - 1 - o.target, cm, o.kind == 1 ? X : reinterpret_cast<const double *>(o.m));
  }
  int T, R;
  effective_tile(opt, nlocal, T, R);
  if (T == 0) return -2;
  opt.tile_bits = T;
  opt.reg_bits = R;
  if (opt.oop && opt.low_bits < opt.oop_low_bits) opt.low_bits = std::min(opt.oop_low_bits, T);
  std::vector<PhysOp> pops;
  for (const auto &h : q.ops) {
    if (h.dead) continue;
    PhysOp p;
    p.type = h.type;
    p.target = h.target;
    p.ctrl = h.ctrl;
    std::memcpy(p.m, h.m, sizeof(h.m));
    pops.push_back(p);
  }
  PlanResult plan = plan_passes(pops, nlocal, 1, opt, nullptr);
  if (plan.passes.empty()) return -3;
  PassPlan &pp = plan.passes.back();
  DevPass &LP = *reinterpret_cast<DevPass *>(pp.blob.data());
  const std::vector<SwapPair> sw = {{nlocal, lbit0}, {nlocal + 1, lbit1}};
  if (!fused_exchange_geometry(LP, nlocal, 1, 4, sw, &LP.xch)) return -4;
  JitProgram dp, plain;
  std::string why;
  if (getenv("QBE_XCH_PLAIN")) LP.xch.n = 0;  // (the same pass without the swap, to compare the generated code)
  if (!jit_generate(pp, JIT_DEVICE_SRC, dp, &why)) return -6;
  if ((int64_t)dp.src.size() + 1 > cap) return -7;
  std::memcpy(out, dp.src.c_str(), dp.src.size() + 1);
  if (LP.xch.n == 0) return 0;
  // the structure key tells the two kinds of pass apart, and nothing of the geometry leaks into it
  JitProgram k1, k2;
  if (!jit_quick(pp, k1, nullptr)) return -8;
  LP.xch.lbit[0] ^= 1u;
  LP.xch.rbase ^= 2u;
  if (!jit_quick(pp, k2, nullptr) || k1.key != k2.key) return -9;
  LP.xch.n = 0;
  if (!jit_quick(pp, plain, nullptr) || plain.key == k1.key) return -10;
  return 0;
}

}  // extern "C"
