// Structure-specialised fused passes: generator (host only) + NVRTC runtime.  See qb_jit.h.
#include "qb_jit.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <array>
#include <sstream>

namespace qb {

static const char kPrelude[] =
#include "qb_jit_prelude.inc"
    ;
const char *jit_prelude() { return kPrelude; }

namespace {

// Every literal that shapes the generated code goes through lit(): it is appended to the
// structural key whether or not source text is being produced, so equal keys mean equal source.
struct Gen {
  bool want_src = false, host = false;
  std::string key;
  std::ostringstream o;
  std::ostringstream *cur = &o;  // where line() / barrier() write
  std::vector<double> coefs;
  double left_out = 1.0;

  void k64(uint64_t v) { key.append(reinterpret_cast<const char *>(&v), sizeof(v)); }
  std::string lit(uint64_t v, const char *suffix = "u") {
    k64(v);
    if (!want_src) return std::string();
    char b[40];
    std::snprintf(b, sizeof(b), "0x%llx%s", (unsigned long long)v, suffix);
    return b;
  }
  std::string dec(int64_t v) {
    k64((uint64_t)v);
    if (!want_src) return std::string();
    return std::to_string(v);
  }
  void tag(const char *t) {  // a structural choice without a number (which helper is called)
    key.append(t);
    key.push_back('\0');
  }
  int coef(double v) {
    coefs.push_back(v);
    return (int)coefs.size() - 1;
  }
  void line(const std::string &s) {
    if (want_src) *cur << "    " << s << "\n";
  }
  // a barrier between thread phases: the host emulation runs the threads of a CTA one after the
  // other, phase by phase
  void barrier(bool warp_only) {
    tag(warp_only ? "bw" : "bc");
    if (!want_src) return;
    if (host) *cur << "  }\n  for (u32 tid = 0; tid < QBJ_NT; ++tid) {\n    QBJ_THREAD_REFS\n";
    else *cur << (warp_only ? "    __syncwarp();\n" : "    QBJ_BAR();\n");
  }
};

std::string reg_offset_expr(Gen &g, const uint8_t *pos, const DevRound &rd, int R, int i, int from_bit, const char *op) {
  // XOR / sum of the global strides of the set register bits of i (bits >= from_bit); pos = where
  // each tile-local bit lives (DevPass::tile_pos for loads, DevPass::out_pos for stores)
  uint64_t off = 0;
  for (int j = from_bit; j < R; ++j)
    if ((i >> j) & 1) off |= 1ull << pos[rd.reg_pos[j]];
  (void)op;
  return g.lit(off, "ull");
}

}  // namespace

bool jit_generate(const PassPlan &pp, JitEmit emit, JitProgram &out, std::string *why) {
  auto fail = [&](const char *m) {
    if (why) *why = m;
    return false;
  };
  if (pp.blob.size() < sizeof(DevPass)) return fail("short blob");
  const DevPass &P = *reinterpret_cast<const DevPass *>(pp.blob.data());
  if (P.lite == 0) return fail("not a step (lite) pass");
  // profiling switches (results are wrong by construction): honoured here only with bit 4 set --
  // bit 0 no global loads, 1 no global stores, 2 no transposes, 3 no gates (plain tile loop only)
  if (P.dbg_skip && !(P.dbg_skip & 16u)) return fail("profiling switches set");
  const uint32_t dbg = P.dbg_skip & 15u;
  if (dbg && (P.tma != 0 || P.jit_group > 1)) return fail("profiling switches: plain tile loop only");
  if (pp.blob.size() < sizeof(DevPass) + size_t(P.nsteps) * sizeof(DevStep)) return fail("short blob");
  const DevStep *S = reinterpret_cast<const DevStep *>(pp.blob.data() + sizeof(DevPass));
  const int T = (int)P.tile_bits, R = (int)P.reg_bits, NT = 1 << (T - R), NR = 1 << R;
  const int nrounds = (int)P.nrounds;
  if (R < 3 || R > kMaxRegBits || T > kMaxTileBits || T - R > 10 || nrounds < 1 || nrounds > kMaxRounds)
    return fail("unsupported geometry");
  if (!(P.tile_pos[0] == 0 && P.tile_pos[1] == 1 && P.tile_pos[2] == 2)) return fail("tile without the low line bits");

  Gen g;
  g.want_src = emit != JIT_KEY_ONLY;
  g.host = emit == JIT_HOST_SRC;
  g.tag("qbj1");
  g.dec(dbg);
  const int l2pf = (int)P.l2_prefetch;
  // contiguous low part of the tile (a "chunk": 2^cbits amplitudes) and how many consecutive tile
  // ids are neighbours in memory (the lowest run of free bits starts right above the chunk)
  int cbits = 3;
  while (cbits < T && P.tile_pos[cbits] == cbits) ++cbits;
  int group = 1;
  if (l2pf > 0 && P.nruns > 0 && (int)P.run_shift[0] == cbits) {
    const int want = P.jit_group ? (int)P.jit_group : 1;
    while (group * 2 <= want && group * 2 <= (1 << std::min<uint32_t>(P.run_len[0], 8))) group *= 2;
  }
  // Bulk-asynchronous tile load (option "tma"): every 128-byte line of the NEXT tile is copied
  // global -> shared by the copy engine (cp.async.bulk, completion counted on an mbarrier) into
  // the transpose buffer as soon as the last transpose of the running tile has been read back, i.e.
  // while the last round's gates and the stores of the running tile execute; round 0 then takes
  // its registers from shared memory (tile-local index u at byte 16 u).  No load holds registers
  // or scoreboard slots while DRAM answers, and the LSU sees 128-byte shared wavefronts instead of
  // 64-byte global ones.  Needs tile bits 0..2 on the three lowest lanes in round 0 (one whole
  // line per quarter warp: conflict-free without a swizzle, which a plain bulk copy cannot apply).
  bool tma = P.tma != 0 && T - R >= 3 && l2pf <= 1;
  if (tma) {
    uint32_t low = 0;
    for (int k = 0; k < 3; ++k) low |= 1u << P.rounds[0].tid_pos[k];
    if (low != 7u) tma = false;
  }
  if (tma) group = 1;
  const bool oop = P.oop != 0;  // tiles are stored as contiguous blocks of the destination, bits permuted (out_pos)
  if (oop) group = 1;
  const bool xch = oop && P.xch.n != 0;  // the stores carry a global<->local swap: every tile goes to the rank that owns it next
  const size_t tables_bytes = size_t(std::max(1, 2 * (nrounds - 1))) * NT * sizeof(uint16_t) + size_t(2) * NT * sizeof(uint64_t) +
                              (size_t(group) << (T - 3)) * sizeof(uint32_t) + (tma ? NT * sizeof(uint16_t) + 16 : 0);
  // tma = 2: ONE CTA per SM holds TWO groups of 2^(T-R) threads, each working on its own tile with a
  // transpose buffer of its own, plus ONE input buffer the copy engine fills: the group that has
  // just taken its registers out of the input buffer starts the copy of the next tile -- which
  // belongs to the OTHER group -- so every group's next tile arrives while it still computes the
  // running one (a true double buffer: 3 x 2^T x 16 bytes of shared memory, 192 KB at T = 12).
  const bool dual = tma && P.tma >= 2 && 3 * (size_t(16) << T) + tables_bytes + 64 <= 227u * 1024u && 2 * NT <= 1024;
  const size_t smem = (dual ? 3 : 1) * (size_t(16) << T) + tables_bytes;
  int minb;
  {
    const int regs_wanted = 4 * NR + 64;
    const int by_regs = 65536 / (NT * regs_wanted);
    const int by_smem = (int)((227u * 1024u) / (smem + 1024));
    minb = std::max(1, std::min(std::min(by_regs, by_smem), 8));
    if (P.jit_minb) minb = std::max(1, std::min((int)P.jit_minb, by_smem));
    if (dual) minb = 1;
  }
  const std::string sT = g.dec(T), sR = g.dec(R), sNT = g.dec(NT), sMINB = g.dec(minb), sNROUNDS = g.dec(nrounds);
  const std::string sL2 = g.dec(l2pf);
  const bool has_gs = P.has_gscale != 0;
  g.dec(has_gs);
  const int LPT = 1 << (R - 3);
  const int npf = (P.pf_lines > 0 && (int)P.pf_lines < LPT) ? (int)P.pf_lines : LPT;  // lines per thread the L2 prefetch covers
  g.dec(npf);
  // ---- one shared-memory swizzle PER TRANSPOSE.  The slot of tile-local index u is
  // u ^ (XOR over the set bits p >= 3 of u of col[p]), col[p] in 1..7: a bijection for any choice.
  // A 128-bit access of a quarter warp is conflict-free iff the three lowest lane bits move the
  // 16-byte bank group independently: the vectors v(p) = (p < 3 ? 1 << p : col[p]) of their tile
  // positions must be linearly independent over GF(2) -- on the store side (previous round's
  // layout) AND on the load side (this round's).  The fixed swizzle of the generic kernels
  // (col[p] = 1 << (p % 3)) leaves 2-way conflicts where the planner kept lane bits whose
  // positions share a residue (ncu: ~1/3 of all shared-store wavefronts); here the columns of
  // the six positions involved are simply searched.
  const int ntab = std::max(1, 2 * (nrounds - 1));
  std::vector<std::array<uint8_t, 16>> swz_col(nrounds);
  int swz_conflicts = 0, swz_fixed = 0;
  for (int r = 1; r < nrounds; ++r) {
    std::array<uint8_t, 16> col{};
    for (int p = 0; p < 16; ++p) col[p] = (uint8_t)(1u << (p % 3));
    swz_col[r] = col;
    if (T - R < 3) continue;
    const DevRound &PR = P.rounds[r - 1], &RD = P.rounds[r];
    // the flip-mask toggles of the previous round: a pending flip XORs the register bit's
    // contribution into the STORE address of the threads that hold it, and which threads do
    // depends on thread-id bits -- possibly on the three lowest lane bits
    struct Tg { uint32_t cthr, bit; bool ext; };
    std::vector<Tg> togs;
    for (uint32_t si = PR.step_begin; si < PR.step_end; ++si)
      for (uint32_t k = 0; k < S[si].ntog && k < (uint32_t)kStepToggles; ++k)
        if (S[si].tog[k].cthr || S[si].tog[k].cext) togs.push_back({S[si].tog[k].cthr, S[si].tog[k].bit, S[si].tog[k].cext != 0});
    auto vec = [&](int pos) { return pos < 3 ? (uint8_t)(1u << pos) : col[pos]; };
    auto collisions = [&]() {
      int bad = 0;
      for (int scen = 0; scen < 4; ++scen) {  // the other thread-id bits all 0 / all 1, external controls unmet / met
        const uint32_t others = (scen & 1) ? (uint32_t)(NT - 1) & ~7u : 0u;
        const bool ext_ok = (scen & 2) != 0;
        uint32_t seen_st = 0, seen_ld = 0;
        for (uint32_t l = 0; l < 8; ++l) {
          const uint32_t tid = l | others;
          uint32_t f = 0;
          for (const Tg &t : togs)
            if ((tid & t.cthr) == t.cthr && (!t.ext || ext_ok)) f ^= 1u << t.bit;
          uint8_t a = 0, b = 0;
          for (int k = 0; k < 3; ++k)
            if ((l >> k) & 1u) {
              a ^= vec(PR.tid_pos[k]);
              b ^= vec(RD.tid_pos[k]);
            }
          for (int j = 0; j < R; ++j)
            if ((f >> j) & 1u) a ^= vec(PR.reg_pos[j]);
          if (seen_st & (1u << a)) ++bad;
          if (seen_ld & (1u << b)) ++bad;
          seen_st |= 1u << a;
          seen_ld |= 1u << b;
        }
      }
      return bad;
    };
    int best = collisions();
    if (best > 0) {
      std::vector<int> freep;  // positions >= 3 whose column matters: the six lane positions, the toggled register bits
      auto add = [&](int pos) {
        if (pos >= 3 && std::find(freep.begin(), freep.end(), pos) == freep.end()) freep.push_back(pos);
      };
      for (int k = 0; k < 3; ++k) {
        add(PR.tid_pos[k]);
        add(RD.tid_pos[k]);
      }
      for (const Tg &t : togs) add(PR.reg_pos[t.bit]);
      std::array<uint8_t, 16> bestcol = col;
      uint64_t rng = 0x9E3779B97F4A7C15ull + (uint64_t)r;
      for (int tries = 0; tries < 4000 && best > 0; ++tries) {
        for (int pos : freep) {
          rng = rng * 6364136223846793005ull + 1442695040888963407ull;
          col[pos] = (uint8_t)(1 + (rng >> 33) % 7);
        }
        const int c = collisions();
        if (c < best) {
          best = c;
          bestcol = col;
        }
      }
      col = bestcol;
      if (best > 0) ++swz_conflicts; else ++swz_fixed;
    }
    swz_col[r] = col;
  }
  auto swz_of = [&](int r, uint32_t u) {  // swizzle of transpose r (linear over XOR)
    uint32_t x = u;
    for (int p = 3; p < T; ++p)
      if ((u >> p) & 1u) x ^= swz_col[r][p];
    return x;
  };
  const std::string sG = g.dec(group), sCB = g.dec(cbits), sMEM = g.dec(P.jit_mem);
  g.dec(tma);
  g.dec(dual);
  g.dec(oop);
  g.dec(xch);
  const std::string sPFK = g.dec((group > 1 && P.jit_pf_last) ? group - 1 : 0);  // prefetch while this tile of the group computes

  // ---------------------------------------------------------------- fragments shared by both modes
  auto tid_bits_expr = [&](const DevRound &rd, int phys, const char *var) {
    // OR of ((var >> j) & 1) << position, positions as literals (0 tile-local, 1 where the tile is
    // loaded from, 2 where it is stored)
    std::string e;
    for (int j = 0; j < T - R; ++j) {
      const uint64_t pos = phys == 2 ? P.out_pos[rd.tid_pos[j]] : (phys ? P.tile_pos[rd.tid_pos[j]] : rd.tid_pos[j]);
      const std::string sp = g.dec((int64_t)pos);
      if (!g.want_src) continue;
      if (!e.empty()) e += " | ";
      e += phys ? "((u64)((" : "(((";
      e += var;
      e += " >> " + std::to_string(j) + ") & 1u) << " + sp + ")";
    }
    return e;
  };
  auto deposit = [&](const char *dst, const char *id) {
    // dst = tile id scattered into the free non-tile bit runs | A.base_fixed
    g.line(std::string("{ u64 t_ = ") + id + "; " + dst + " = A.base_fixed;");
    for (uint32_t k = 0; k < P.nruns; ++k) {
      const std::string len = g.dec(P.run_len[k]), sh = g.dec(P.run_shift[k]);
      g.line(std::string("  ") + dst + " |= (t_ & ((1ull << " + len + ") - 1ull)) << " + sh + "; t_ >>= " + len + ";");
    }
    g.line("}");
  };
  const DevRound &R0 = P.rounds[0];
  const DevRound &RL = P.rounds[nrounds - 1];
  auto stride_of = [&](const DevRound &rd, int j) { return 1ull << P.tile_pos[rd.reg_pos[j]]; };
  auto ostride_of = [&](const DevRound &rd, int j) { return 1ull << P.out_pos[rd.reg_pos[j]]; };
  const std::string sOB = oop ? "obase_" : "base";  // (out of place: set by every tile loop before its store)
  // out of place: the block address of tile number `id` (its bits, group by group, at their new places)
  auto emit_obase = [&](std::ostream &os, const char *id, const char *indent) {
    os << indent << "u64 obase_ = 0;\n" << indent << "{ u64 t_ = " << id << ";\n";
    for (uint32_t k = 0; k < P.onruns && k < (uint32_t)kMaxOutRuns; ++k) {
      const std::string len = g.dec(P.orun_len[k]), sh = g.dec(P.orun_shift[k]);
      os << indent << "  obase_ |= (t_ & ((1ull << " << len << ") - 1ull)) << " << sh << "; t_ >>= " << len << ";\n";
    }
    os << indent << "}\n";
  };

  auto emit_tables = [&]() {
    // per transpose: this thread's swizzled slot in the STORE layout (table 2(r-1)) and in the
    // LOAD layout (table 2(r-1)+1); the swizzle is linear, so each thread-id bit XORs a literal
    for (int r = 1; r < nrounds; ++r)
      for (int side = 0; side < 2; ++side) {
        const DevRound &rd = P.rounds[side ? r : r - 1];
        std::string e = "0u";
        for (int j = 0; j < T - R; ++j)
          e += " ^ ((0u - ((tid >> " + std::to_string(j) + ") & 1u)) & " + g.lit(swz_of(r, 1u << rd.tid_pos[j])) + ")";
        g.line("sidx_tab[" + std::to_string(2 * (r - 1) + side) + " * QBJ_NT + tid] = (u16)(" + e + ");");
      }
    if (tma) g.line("lin_tab[tid] = (u16)(" + tid_bits_expr(R0, 0, "tid") + ");");
    g.line("goff_tab[tid] = " + tid_bits_expr(R0, 1, "tid") + ";");
    g.line("goff_tab[QBJ_NT + tid] = " + tid_bits_expr(RL, 2, "tid") + ";");
    for (int k = 0; k < LPT; ++k) {
      std::string e;
      for (int j = 0; j < T - 3; ++j) {
        const std::string sp = g.dec(P.tile_pos[3 + j]);
        if (!g.want_src) continue;
        if (!e.empty()) e += " | ";
        e += "((u64)(((tid + " + std::to_string(k) + "u * QBJ_NT) >> " + std::to_string(j) + ") & 1u) << " + sp + ")";
      }
      g.line("line_tab[" + std::to_string(k) + " * QBJ_NT + tid] = (u32)((" + e + ") >> 3);");
    }
  };

  auto emit_load = [&]() {
    g.tag("ld");
    g.line(std::string(dbg & 1u ? "if (dbg_never_) " : "") + "{ const u64 src_ = base + goff_tab[tid];");
    if (stride_of(R0, 0) == 1ull) {
      g.tag("p");
      for (int i = 0; i < NR; i += 2) g.line("  QBJ_LD2(src_ + " + reg_offset_expr(g, P.tile_pos, R0, R, i, 1, "+") + ", " + std::to_string(i) + ");");
    } else {
      for (int i = 0; i < NR; ++i) g.line("  QBJ_LD1(src_ + " + reg_offset_expr(g, P.tile_pos, R0, R, i, 0, "+") + ", " + std::to_string(i) + ");");
    }
    g.line("}");
  };
  // round 0 takes its registers from the bulk-copied tile: tile-local index u at byte 16 u
  auto emit_load_smem = [&]() {
    g.tag("lds0");
    g.line("{ const u32 ul0_ = (u32)lin_tab[tid] << 4;");
    for (int i = 0; i < NR; ++i) {
      uint32_t u = 0;
      for (int j = 0; j < R; ++j)
        if ((i >> j) & 1) u |= 1u << R0.reg_pos[j];
      g.line("  QBJ_LDSI(ul0_ | " + g.lit(u << 4) + ", " + std::to_string(i) + ");");
    }
    g.line("}");
    g.barrier(false);  // every thread holds its registers: the buffer is free for the transposes
  };
  // mf: register bits on which a flip may be pending in some thread
  auto emit_store = [&](uint32_t mf) {
    g.tag("st");
    g.dec(mf);
    g.line(std::string(dbg & 2u ? "if (dbg_never_) " : "") + "{ u64 fx_ = 0;");
    for (int j = 0; j < R; ++j)
      if ((mf >> j) & 1u)
        g.line("  fx_ |= ((f >> " + std::to_string(j) + ") & 1u) ? " + g.lit(ostride_of(RL, j), "ull") + " : 0ull;");
    if (ostride_of(RL, 0) == 1ull) {
      g.tag("p");
      g.line("  const u64 at_ = (" + sOB + " + goff_tab[QBJ_NT + tid]) ^ (fx_ & ~1ull);");
      if (xch) g.line("  QBJ_XCH_TILE(at_)");
      const std::string st2 = xch ? "  QBJ_ST2X(" : "  QBJ_ST2(at_ ^ ";
      const bool sw = (mf & 1u) != 0;
      if (sw) g.line("  const bool sw_ = (f & 1u) != 0;");
      for (int i = 0; i < NR; i += 2) {
        const std::string a = std::to_string(i), b = std::to_string(i + 1);
        const std::string off = reg_offset_expr(g, P.out_pos, RL, R, i, 1, "^");
        const std::string head = xch ? st2 + a + ", " + off : st2 + off;
        if (sw)
          g.line(head + ", sw_ ? re[" + b + "] : re[" + a + "], sw_ ? im[" + b + "] : im[" + a + "], sw_ ? re[" + a +
                 "] : re[" + b + "], sw_ ? im[" + a + "] : im[" + b + "]);");
        else
          g.line(head + ", re[" + a + "], im[" + a + "], re[" + b + "], im[" + b + "]);");
      }
    } else {
      g.line("  const u64 at_ = (" + sOB + " + goff_tab[QBJ_NT + tid]) ^ fx_;");
      if (xch) g.line("  QBJ_XCH_TILE(at_)");
      for (int i = 0; i < NR; ++i)
        g.line(std::string(xch ? "  QBJ_ST1X(" + std::to_string(i) + ", " : "  QBJ_ST1(at_ ^ ") + reg_offset_expr(g, P.out_pos, RL, R, i, 0, "^") +
               ", re[" + std::to_string(i) + "], im[" + std::to_string(i) + "]);");
    }
    g.line("}");
  };
  auto sx_of = [&](int r, const DevRound &rd, int i) {  // byte offset contribution of register index i in transpose r
    uint32_t u = 0;
    for (int j = 0; j < R; ++j)
      if ((i >> j) & 1) u |= 1u << rd.reg_pos[j];
    return swz_of(r, u) << 4;
  };

  // ---------------------------------------------------------------- the rounds (both modes)
  uint32_t mf_end = 0;  // flips possibly pending when the last round ends
  bool bad = false;
  // tma: the transpose buffer is free from here to the end of the tile -- start the NEXT tile's bulk copies
  auto issue_next = [&](const char *bar = nullptr) {
    g.tag("tmaissue");
    if (g.host) return;  // (the host emulation copies the tile at the top of its tile loop)
    const std::string b = bar ? bar : (dual ? "(bar_a_ + 8u * (grp_ ^ 1u))" : "bar_a_");
    g.line("if (have_next_) {");
    g.line("  qbj_fence_proxy_async();  // generic-proxy reads of the buffer happen before the async-proxy writes");
    g.line("  if (tid == 0) qbj_mbar_expect_tx(" + b + ", 16u << QBJ_T);");
    // one copy per contiguous CHUNK of the tile (2^cbits amplitudes; a line = 8): chunk c lands at byte
    // c << (cbits + 4) of the buffer and comes from where line c << (cbits - 3) of the tile lives
    const int nchunks = 1 << (T - cbits), cpt = std::max(1, nchunks / NT);
    const std::string sBytes = g.lit(16u << cbits), sCb = g.dec(cbits);
    for (int k = 0; k < cpt; ++k) {
      const std::string c = "(tid + " + std::to_string(k) + "u * QBJ_NT)";
      std::string ln = "  ";
      if (nchunks < NT) ln += "if (tid < " + std::to_string(nchunks) + "u) ";
      ln += "qbj_bulk_load(in_a_ + (" + c + " << (" + sCb + " + 4)), src + next_base + ((u64)line_tab[" + c + " << (" + sCb + " - 3)] << 3), " + sBytes + ", " + b + ");";
      g.line(ln);
    }
    g.line("}");
  };
  auto emit_rounds = [&]() {
    uint32_t mf = 0;
    if (tma && (nrounds == 1 || dual)) issue_next();
    for (int r = 0; r < nrounds; ++r) {
      const DevRound &RD = P.rounds[r];
      if (r > 0 && (dbg & 4u)) {
        g.line("f = 0;");
        mf = 0;
      } else if (r > 0) {
        const DevRound &PR = P.rounds[r - 1];
        const bool local = RD.warp_local != 0;
        const std::string sr = std::to_string(r);
        g.tag("tr");
        g.dec(mf);
        if (local) g.barrier(true);  // (the lanes of this warp finished reading the previous layout)
        g.line("u32 us" + sr + " = (u32)sidx_tab[" + std::to_string(2 * (r - 1)) + " * QBJ_NT + tid] << 4;");
        for (int j = 0; j < R; ++j)
          if ((mf >> j) & 1u)
            g.line("us" + sr + " ^= (0u - ((f >> " + std::to_string(j) + ") & 1u)) & " + g.lit(swz_of(r, 1u << PR.reg_pos[j]) << 4) + ";");
        g.line("f = 0;");
        mf = 0;
        for (int i = 0; i < NR; ++i)
          g.line("QBJ_STS(us" + sr + " ^ " + g.lit(sx_of(r, PR, i)) + ", re[" + std::to_string(i) + "], im[" + std::to_string(i) + "]);");
        g.barrier(local);
        g.line("const u32 ul" + sr + " = (u32)sidx_tab[" + std::to_string(2 * (r - 1) + 1) + " * QBJ_NT + tid] << 4;");
        for (int i = 0; i < NR; ++i) g.line("QBJ_LDS(ul" + sr + " ^ " + g.lit(sx_of(r, RD, i)) + ", " + std::to_string(i) + ");");
        // free the buffer for the next CTA-wide transpose (of this tile, or the first of the next)
        // (wrapping to the next tile: also when the warps' slot regions differ between the last and
        //  the first round, see qb_planner.cpp on rounds[0].warp_local)
        if (dual && r + 1 == nrounds) {
          // (nothing: the group barrier after the next tile's register fill separates this tile's
          //  transposes from the next one's, and the input buffer is not this buffer)
          g.tag("dualend");
        } else if (tma && r + 1 == nrounds) {  // the copy engine overwrites the whole buffer next: everybody must be done with it
          g.barrier(false);
          issue_next();
        } else if (r + 1 < nrounds ? P.rounds[r + 1].warp_local == 0 : (P.rounds[1].warp_local == 0 || P.rounds[0].warp_local == 0))
          g.barrier(false);
      }
      for (uint32_t si = RD.step_begin; si < RD.step_end; ++si) {
        const DevStep &st = S[si];
        if (dbg & 8u) continue;
        for (int J = 0; J < R; ++J) {
          const uint32_t kind = (st.kinds >> (4 * J)) & 15u;
          const uint32_t cls = kind & SLOT_CLASS;
          if (cls == SLOT_NONE) continue;
          const bool flip = ((mf >> J) & 1u) != 0;
          if (flip && !(kind & SLOT_FLIP)) bad = true;  // the planner says no flip can be pending here
          const std::string sJ = std::to_string(J);
          if (cls == SLOT_ROT) {
            const double cs = st.slot[J][2], sn = st.slot[J][3];
            if (!(cs >= 0.0) || !std::isfinite(cs) || !std::isfinite(sn) || !(cs * cs + sn * sn > 0.5)) bad = true;
            const bool formA = std::fabs(sn) <= cs;
            if (formA) {
              const int k = g.coef(sn / cs);
              g.left_out *= cs;
              if (flip) {
                g.tag("raf");
                g.line("qbj_rot_a_flip<" + g.dec(J) + ">(re, im, QBJ_C(" + g.dec(k) + "), f);");
              } else {
                g.tag("ra");
                g.line("qbj_rot_a<" + g.dec(J) + ">(re, im, QBJ_C(" + g.dec(k) + "));");
              }
            } else if (!flip) {
              const int k = g.coef(cs / sn);
              g.left_out *= sn;
              g.tag("rb");
              g.line("qbj_rot_b<" + g.dec(J) + ">(re, im, QBJ_C(" + g.dec(k) + "));");
            } else {
              const int k = g.coef(st.slot[J][0]);
              g.coef(st.slot[J][1]);
              g.tag("r3f");
              g.line("qbj_rot3_flip<" + g.dec(J) + ">(re, im, QBJ_C(" + g.dec(k) + "), QBJ_C(" + std::to_string(k + 1) + "), f);");
            }
          } else if (cls == SLOT_REAL) {
            const int k = g.coef(st.slot[J][0]);
            for (int e = 1; e < 4; ++e) g.coef(st.slot[J][e]);
            g.tag("re");
            g.line("qbj_real<" + g.dec(J) + ", " + g.dec(flip ? 1 : 0) + ">(re, im, &QBJ_C(" + g.dec(k) + "), f);");
          } else if (cls == SLOT_GENERAL || cls == SLOT_GENERAL1) {
            const int k = g.coef(st.slot[J][0]);
            for (int e = 1; e < 8; ++e) g.coef(st.slot[J][e]);
            if (cls == SLOT_GENERAL1 && !flip) {
              g.tag("g1");
              g.line("qbj_general1<" + g.dec(J) + ">(re, im, &QBJ_C(" + g.dec(k) + "));");
            } else {
              g.tag("ge");
              g.line("qbj_general<" + g.dec(J) + ", " + g.dec(flip ? 1 : 0) + ">(re, im, &QBJ_C(" + g.dec(k) + "), f);");
            }
          } else {
            bad = true;
          }
          (void)sJ;
        }
        if (st.ntog > (uint32_t)kStepToggles) bad = true;
        for (uint32_t k = 0; k < st.ntog && k < (uint32_t)kStepToggles; ++k) {
          const auto &tg = st.tog[k];
          if ((int)tg.bit >= R) bad = true;
          if (tg.cthr == 0 && tg.cext == 0) {  // uncontrolled X: a renaming
            g.tag("xa");
            g.line("qbj_swap_all<" + g.dec(tg.bit) + ">(re, im);");
            continue;
          }
          g.tag("tg");
          std::string cond;
          const std::string sth = g.lit(tg.cthr), sex = g.lit(tg.cext, "ull");
          if (tg.cthr) cond = "((tid & " + sth + ") == " + sth + ")";
          if (tg.cext) cond += std::string(cond.empty() ? "" : " && ") + "((basefull & " + sex + ") == " + sex + ")";
          g.line("f ^= (" + cond + ") ? " + g.lit(1u << tg.bit) + " : 0u;");
          mf |= 1u << tg.bit;
        }
        if (st.swap_j != 0xffu) {
          const int J = (int)(st.swap_j & 7u);
          if (J >= R || st.swap_creg == 0 || (st.swap_creg >> R) != 0 || ((st.swap_creg >> J) & 1u)) bad = true;
          const bool stat = __builtin_popcount(st.swap_creg) == 1 && st.swap_cthr == 0 && st.swap_cext == 0 && !(mf & st.swap_creg);
          if (stat) {
            g.tag("ss");
            g.line("qbj_swap_static<" + g.dec(J) + ", " + g.dec(__builtin_ctz(st.swap_creg)) + ">(re, im);");
          } else {
            g.tag("sd");
            std::string cond = "true";
            const std::string sth = g.lit(st.swap_cthr), sex = g.lit(st.swap_cext, "ull");
            if (st.swap_cthr) cond = "((tid & " + sth + ") == " + sth + ")";
            if (st.swap_cext) cond += " && ((basefull & " + sex + ") == " + sex + ")";
            g.line("qbj_swap_dyn<" + g.dec(J) + ">(re, im, " + g.lit(st.swap_creg) + ", " + cond + ", f);");
          }
        }
      }
    }
    mf_end = mf;
    if (has_gs) g.line("qbj_scale(re, im, A.gs[0], A.gs[1]);");
  };

  // ---------------------------------------------------------------- assemble
  if (!g.want_src) {
    // key only: walk every fragment once (order irrelevant as long as it is fixed)
    emit_tables();
    deposit("base", "tile_id");
    if (tma) emit_load_smem(); else emit_load();
    emit_rounds();
    emit_store(mf_end);
    g.dec(P.onruns);
    for (uint32_t k = 0; k < P.onruns && k < (uint32_t)kMaxOutRuns; ++k) {
      g.dec(P.orun_len[k]);
      g.dec(P.orun_shift[k]);
    }
  } else {
    std::ostringstream &o = g.o;
    o << "#define QBJ_T " << sT << "\n#define QBJ_R " << sR << "\n#define QBJ_NT " << sNT << "\n#define QBJ_MEM " << sMEM << "\n";
    if (g.host) o << "#define QB_JIT_HOST 1\n";
    o << kPrelude << "\n";
    // the rounds are generated first (into a side buffer): the coefficient count sizes QbjArgs
    std::ostringstream side;
    g.cur = &side;
    emit_rounds();
    g.cur = &g.o;
    const std::string rounds_txt = side.str();
    const size_t nc = std::max<size_t>(1, g.coefs.size());
    o << "struct QbjXch { u64 peer[" << kMaxXchRanks << "]; u64 vmask, vconst; u32 n, rbase; u32 lbit[4], rbit[4]; unsigned char dr[32]; u32 pad_[2]; };\n"
         "struct QbjArgs { double gs[2]; u64 rank_bits; u64 base_fixed; QbjXch x; double c[" << nc << "]; };\n";
    if (xch) {
      // the stores carry a global<->local swap (XchGeom): amplitude address a of the new layout goes to the
      // rank its victims' bits pick, at the address with those bits replaced by this rank's old rank-bit
      // values.  Every store of a thread is its base address `at` XOR a literal offset, so the run-time bit
      // positions are looked at ONCE per tile (rank and address of the base); a store then XORs what its
      // register index adds to the rank (dr[i], worked out on the host) and its offset outside the victims.
      g.tag("xch");
      o << "#define QBJ_XCH_TILE(at) u32 rrt_ = A.x.rbase; \\\n"
           "  for (u32 k_ = 0; k_ < A.x.n; ++k_) rrt_ |= (u32)(((at) >> A.x.lbit[k_]) & 1ull) << A.x.rbit[k_]; \\\n"
           "  const u64 nvm_ = ~A.x.vmask; const u64 bt_ = ((at) & nvm_) | A.x.vconst;\n";
      if (g.host)
        o << "#define QBJ_DSTX(i, off) (reinterpret_cast<double *>(A.x.peer[rrt_ ^ A.x.dr[i]]) + 2 * (bt_ ^ ((off) & nvm_)))\n";
      else
        o << "#define QBJ_DSTX(i, off) (reinterpret_cast<double2 *>(A.x.peer[rrt_ ^ A.x.dr[i]]) + (bt_ ^ ((off) & nvm_)))\n";
    }
    // Coefficients are loop-invariant kernel parameters.  Up to ~30 of them the compiler keeps them
    // in uniform registers across the tile loop; beyond that it hoists them into ordinary
    // registers and spills those (the general-class passes of the benchmark: 200-300 bytes per
    // thread re-read through L1 for every tile).  There the index gets a zero the compiler cannot
    // see through and that is re-made every tile (tile_id >> 31), so each use is one indexed
    // constant-bank load and nothing is hoisted.
    const bool coef_reload = !g.host && g.coefs.size() > 30;  // (63 uniform registers = 31 doubles)
    g.dec(coef_reload);
    o << (coef_reload ? "#define QBJ_C(k) A.c[(k) + cz_]\n" : "#define QBJ_C(k) A.c[k]\n");
    if (!g.host) {
      o << "#define QBJ_LD2(p, i) qbj_ld256(src + (p), re[i], im[i], re[(i) + 1], im[(i) + 1])\n"
           "#define QBJ_LD1(p, i) qbj_ld128(src + (p), re[i], im[i])\n"
           "#define QBJ_ST2(p, a0, a1, b0, b1) qbj_st256(QBJ_DST(p), a0, a1, b0, b1)\n"
           "#define QBJ_ST1(p, xr, xi) qbj_st128(QBJ_DST(p), xr, xi)\n"
           "#define QBJ_STS(off, xr, xi) *reinterpret_cast<double2 *>(sm_ + (off)) = make_double2(xr, xi)\n"
           "#define QBJ_LDS(off, i) { const double2 a_ = *reinterpret_cast<const double2 *>(sm_ + (off)); re[i] = a_.x; im[i] = a_.y; }\n"
           "#define QBJ_LDSI(off, i) { const double2 a_ = *reinterpret_cast<const double2 *>(in_ + (off)); re[i] = a_.x; im[i] = a_.y; }\n";
      o << "#define QBJ_DST(p) (amps + (p))\n";
      if (xch)
        o << "#define QBJ_ST2X(i, off, a0, a1, b0, b1) qbj_st256(QBJ_DSTX(i, off), a0, a1, b0, b1)\n"
             "#define QBJ_ST1X(i, off, xr, xi) qbj_st128(QBJ_DSTX(i, off), xr, xi)\n";
      if (dual)
        o << "#define QBJ_BAR() asm volatile(\"bar.sync %0, %1;\" ::\"r\"(grp_ + 1u), \"r\"((u32)QBJ_NT) : \"memory\")\n";
      else
        o << "#define QBJ_BAR() __syncthreads()\n";
      o << "extern \"C\" __global__ void __launch_bounds__(" << (dual ? "2 * QBJ_NT" : "QBJ_NT") << ", " << sMINB
        << ") qb_jit_pass(double2 *amps, const double2 *src, u64 ntiles, const __grid_constant__ QbjArgs A) {\n"
           "  extern __shared__ __align__(16) unsigned char smem_raw[];\n";
      if (dual)
        o << "  const u32 grp_ = threadIdx.x >> (QBJ_T - QBJ_R), tid = threadIdx.x & (QBJ_NT - 1u);\n"
             "  unsigned char *const sm_ = smem_raw + ((size_t)grp_ << (QBJ_T + 4));\n"
             "  const unsigned char *const in_ = smem_raw + ((size_t)2 << (QBJ_T + 4));\n"
             "  u16 *sidx_tab = reinterpret_cast<u16 *>(smem_raw + ((size_t)3 << (QBJ_T + 4)));\n";
      else
        o << "  const u32 tid = threadIdx.x;\n"
             "  unsigned char *const sm_ = smem_raw;\n"
             "  const unsigned char *const in_ = smem_raw;\n  (void)in_;\n"
             "  u16 *sidx_tab = reinterpret_cast<u16 *>(smem_raw + (16u << QBJ_T));\n";
      o << ""
           "  u64 *goff_tab = reinterpret_cast<u64 *>(sidx_tab + "
        << ntab
        << " * QBJ_NT);\n"
           "  u32 *line_tab = reinterpret_cast<u32 *>(goff_tab + 2 * QBJ_NT);\n";
      if (tma)
        o << "  u16 *lin_tab = reinterpret_cast<u16 *>(line_tab + (1 << (QBJ_T - 3)));\n"
             "  const u32 in_a_ = (u32)__cvta_generic_to_shared(in_);\n"
             "  const u32 bar_a_ = (u32)__cvta_generic_to_shared(lin_tab + QBJ_NT);\n";
      o << "  {\n";
      emit_tables();
      o << "  }\n";
      if (dual) {
      // ---- two groups, one input buffer: group g takes this CTA's tiles k = g, g + 2, ...; copy k + 1 is
      // started by the group that has just emptied the input buffer of copy k
      o << "  if (threadIdx.x == 0) { qbj_mbar_init(bar_a_, 1u); qbj_mbar_init(bar_a_ + 8u, 1u); }\n"
           "  qbj_fence_mbar_init();\n"
           "  __syncthreads();  // (also: the tables above, written by both groups with the same values)\n"
           "  const u32 ntiles32 = (u32)ntiles, stride = gridDim.x, first = blockIdx.x;\n"
           "  const u32 my_iters = first < ntiles32 ? (ntiles32 - first + stride - 1u) / stride : 0u;\n"
           "  double re[QBJ_NR], im[QBJ_NR];\n"
           "  u32 f = 0;\n"
           "  u64 base = 0, next_base = 0;\n"
           "  bool have_next_ = grp_ == 0u && my_iters > 0u;\n"
           "  if (have_next_) {\n";
      deposit("next_base", "first");
      {
        std::ostringstream side2;
        std::ostringstream *keep = g.cur;
        g.cur = &side2;
        issue_next("bar_a_");
        g.cur = keep;
        o << side2.str();
      }
      o << "  }\n";
      if (l2pf > 0) {
        o << "  if (grp_ == 1u && my_iters > 1u) {\n    const u32 pf_id = first + stride;\n    u64 pb_;\n";
        deposit("pb_", "pf_id");
        for (int k = 0; k < LPT; ++k)
          o << "    asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(src + pb_ + ((u64)line_tab[" << k << " * QBJ_NT + tid] << 3)));\n";
        o << "  }\n";
      }
      o << "  for (u32 k_ = grp_; k_ < my_iters; k_ += 2u) {\n"
           "    const u32 tile_id = first + k_ * stride;\n";
      deposit("base", "tile_id");
      o << "    have_next_ = k_ + 1u < my_iters;\n"
           "    if (have_next_) {\n      const u32 next_id = tile_id + stride;\n      u64 nb_;\n";
      deposit("nb_", "next_id");
      o << "      next_base = nb_;\n    }\n";
      if (l2pf > 0) {
        // the tile after the next one goes to L2 now: its bulk copy (issued one tile from now) must
        // not wait for DRAM -- only ONE copy is in flight per SM, its latency is the pipeline's period
        o << "    if (k_ + 2u < my_iters) {\n      const u32 pf_id = tile_id + 2u * stride;\n      u64 pb_;\n";
        deposit("pb_", "pf_id");
        for (int k = 0; k < LPT; ++k)
          o << "      asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(src + pb_ + ((u64)line_tab[" << k << " * QBJ_NT + tid] << 3)));\n";
        o << "    }\n";
      }
      o << "    qbj_mbar_wait(bar_a_ + 8u * grp_, (k_ >> 1) & 1u);\n"
           "    const u64 basefull = base | A.rank_bits;\n    (void)basefull;\n    f = 0;\n";
      if (coef_reload) o << "    const u32 cz_ = tile_id >> 31;  // always 0 (tile ids are < 2^31), but not to the compiler\n";
      {
        std::ostringstream side2;
        std::ostringstream *keep = g.cur;
        g.cur = &side2;
        emit_load_smem();
        g.cur = keep;
        o << side2.str();
      }
      o << rounds_txt;
      if (oop) emit_obase(o, "tile_id", "    ");
      emit_store(mf_end);
      o << "  }\n}\n";
      } else if (tma) {
      // ---- bulk-asynchronous loads: the copy engine fills the transpose buffer one tile ahead
      o << "  if (tid == 0) qbj_mbar_init(bar_a_, 1u);\n"
           "  qbj_fence_mbar_init();\n"
           "  __syncthreads();\n"
           "  const u32 ntiles32 = (u32)ntiles, stride = gridDim.x, first = blockIdx.x;\n"
           "  const u32 iters = (ntiles32 + stride - 1) / stride;\n"
           "  double re[QBJ_NR], im[QBJ_NR];\n"
           "  u32 f = 0, phase_ = 0;\n"
           "  u64 base = 0, next_base = 0;\n"
           "  bool have_next_ = first < ntiles32;\n"
           "  if (have_next_) {\n";
      deposit("next_base", "first");
      {
        std::ostringstream side2;
        std::ostringstream *keep = g.cur;
        g.cur = &side2;
        issue_next();
        g.cur = keep;
        o << side2.str();
      }
      o << "  }\n"
           "  for (u32 it = 0; it < iters; ++it) {\n"
           "    const u32 tile_id = first + it * stride;\n"
           "    if (tile_id >= ntiles32) break;\n"
           "    base = next_base;\n"
           "    { const u32 next_id = tile_id + stride;\n"
           "      have_next_ = next_id < ntiles32;\n"
           "      if (have_next_) {\n      u64 nb_;\n";
      deposit("nb_", "next_id");
      o << "      next_base = nb_;\n";
      if (l2pf > 0)
        for (int k = 0; k < LPT; ++k)
          o << "      asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(src + nb_ + ((u64)line_tab[" << k << " * QBJ_NT + tid] << 3)));\n";
      o << "      }\n    }\n"
           "    qbj_mbar_wait(bar_a_, phase_);\n"
           "    phase_ ^= 1u;\n"
           "    const u64 basefull = base | A.rank_bits;\n    (void)basefull;\n    f = 0;\n";
      if (coef_reload) o << "    const u32 cz_ = tile_id >> 31;  // always 0 (tile ids are < 2^31), but not to the compiler\n";
      {
        std::ostringstream side2;
        std::ostringstream *keep = g.cur;
        g.cur = &side2;
        emit_load_smem();
        g.cur = keep;
        o << side2.str();
      }
      o << rounds_txt;
      if (oop) emit_obase(o, "tile_id", "    ");
      emit_store(mf_end);
      o << "  }\n}\n";
      } else if (group == 1) {
      o << "  const u32 ntiles32 = (u32)ntiles, stride = gridDim.x, first = blockIdx.x;\n"
           "  const u32 iters = (ntiles32 + stride - 1) / stride;\n"
           "  double re[QBJ_NR], im[QBJ_NR];\n"
           "  u32 f = 0;\n"
           "  u64 base = 0, next_base = 0;\n";
      if (dbg)
        o << "  const bool dbg_never_ = ntiles == ~0ull;  // profiling switches: code kept, never executed\n"
             "  for (int i = 0; i < QBJ_NR; ++i) { re[i] = 1e-3 * (double)(tid + i); im[i] = -re[i]; }\n";
      o << "  for (u32 it = 0; it <= iters; ++it) {\n"
           "    const u32 tile_id = first + it * stride;\n"
           "    if (it > 0 && tile_id - stride < ntiles32) {\n";
      if (oop) emit_obase(o, "(tile_id - stride)", "    ");
      // NOTE: the flips pending at the store are those of the previous tile's last round
      // (mf_end was computed by emit_rounds above)
      emit_store(mf_end);
      o << "    }\n"
           "    const bool active = it < iters && tile_id < ntiles32;\n"
           "    if (!active) break;\n";
      if (l2pf == 1) o << "    if (it > 0) base = next_base; else\n";
      deposit("base", "tile_id");
      emit_load();
      if (l2pf > 0) {
        o << "    { const u32 next_id = tile_id + stride * " << sL2 << "u;\n"
             "      if (next_id < ntiles32) {\n      u64 nb_;\n";
        deposit("nb_", "next_id");
        o << "      next_base = nb_;\n";
        for (int k = 0; k < npf; ++k)
          o << "      asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(src + nb_ + ((u64)line_tab[" << k << " * QBJ_NT + tid] << 3)));\n";
        o << "      }\n    }\n";
      }
      o << "    const u64 basefull = base | A.rank_bits;\n    (void)basefull;\n    f = 0;\n";
      if (coef_reload) o << "    const u32 cz_ = tile_id >> 31;  // always 0 (tile ids are < 2^31), but not to the compiler\n";
      o << rounds_txt;
      o << "  }\n}\n";
      } else {
      // ---- grouped order: this CTA takes `group` consecutive tiles in a row (neighbours in memory:
      // together they cover group x chunk contiguous bytes per chunk) and, while it works on a
      // group, pulls the NEXT group into L2 with one bulk request per chunk.
      // group line table: the 128-byte lines of a whole GROUP, enumerated so that consecutive
      // q (= consecutive lanes) walk the contiguous run of group x chunk bytes first:
      //   q = tid + k * NT;  low bits of q = line inside the run, high bits = which chunk
      o << "  {\n";
      const int run_lines_log2 = (cbits - 3) + __builtin_ctz((unsigned)group);  // lines per contiguous run
      const int glpt = group * LPT;                                           // group lines per thread
      for (int k = 0; k < glpt; ++k) {
        std::string e = "(u64)(q_ & " + std::to_string((1 << run_lines_log2) - 1) + "u) << 3";
        for (int j = 0; j < T - cbits; ++j)
          e += " | ((u64)((q_ >> " + std::to_string(run_lines_log2 + j) + ") & 1u) << " + g.dec(P.tile_pos[cbits + j]) + ")";
        o << "    { const u32 q_ = tid + " << k << "u * QBJ_NT; line_tab[" << k << " * QBJ_NT + tid] = (u32)((" << e << ") >> 3); }\n";
      }
      o << "  }\n"
           "  const u32 ntiles32 = (u32)ntiles, stride = gridDim.x;\n"
           "  const u32 ngroups = (ntiles32 + " << sG << "u - 1u) / " << sG << "u;\n"
           "  double re[QBJ_NR], im[QBJ_NR];\n"
           "  u32 f = 0;\n"
           "  u64 base = 0;\n"
           "  u32 grp = blockIdx.x, k_in = 0;\n"
           "  bool have_prev = false;\n"
           "  while (true) {\n"
           "    const u32 tile_id = grp * " << sG << "u + k_in;\n"
           "    if (have_prev) {\n";
      emit_store(mf_end);
      o << "    }\n"
           "    if (!(grp < ngroups && tile_id < ntiles32)) break;\n";
      deposit("base", "tile_id");
      emit_load();
      o << "    if (k_in == " << sPFK << "u) {\n"
           "      const u32 ng_ = grp + stride;\n"
           "      if (ng_ < ngroups) {\n      u64 nb_;\n"
           "      const u32 nt0_ = ng_ * " << sG << "u;\n";
      deposit("nb_", "nt0_");
      for (int k = 0; k < glpt; ++k)
        o << "      asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(src + nb_ + ((u64)line_tab[" << k << " * QBJ_NT + tid] << 3)));\n";
      o << "      }\n    }\n";
      o << "    const u64 basefull = base | A.rank_bits;\n    (void)basefull;\n    f = 0;\n";
      if (coef_reload) o << "    const u32 cz_ = tile_id >> 31;  // always 0 (tile ids are < 2^31), but not to the compiler\n";
      o << rounds_txt;
      o << "    have_prev = true;\n"
           "    if (++k_in == " << sG << "u) { k_in = 0; grp += stride; }\n"
           "  }\n}\n";
      }
    } else {
      o << "#define QBJ_LD2(p, i) { re[i] = amps[2 * (p)]; im[i] = amps[2 * (p) + 1]; re[(i) + 1] = amps[2 * (p) + 2]; im[(i) + 1] = amps[2 * (p) + 3]; }\n"
           "#define QBJ_LD1(p, i) { re[i] = amps[2 * (p)]; im[i] = amps[2 * (p) + 1]; }\n"
           "#define QBJ_ST2(p, a0, a1, b0, b1) { double *d_ = QBJ_DST(p); d_[0] = a0; d_[1] = a1; d_[2] = b0; d_[3] = b1; }\n"
           "#define QBJ_ST1(p, xr, xi) { double *d_ = QBJ_DST(p); d_[0] = xr; d_[1] = xi; }\n"
           "#define QBJ_DST(p) (dst + 2 * (p))\n"
           "#define QBJ_ST2X(i, off, a0, a1, b0, b1) { double *d_ = QBJ_DSTX(i, off); d_[0] = a0; d_[1] = a1; d_[2] = b0; d_[3] = b1; }\n"
           "#define QBJ_ST1X(i, off, xr, xi) { double *d_ = QBJ_DSTX(i, off); d_[0] = xr; d_[1] = xi; }\n"
           "#define QBJ_STS(off, xr, xi) { SM[2 * ((off) >> 4)] = xr; SM[2 * ((off) >> 4) + 1] = xi; }\n"
           "#define QBJ_LDS(off, i) { re[i] = SM[2 * ((off) >> 4)]; im[i] = SM[2 * ((off) >> 4) + 1]; }\n"
           "#define QBJ_LDSI(off, i) QBJ_LDS(off, i)\n"
           "#define QBJ_THREAD_REFS double (&re)[QBJ_NR] = RE[tid]; double (&im)[QBJ_NR] = IM[tid]; u32 &f = F[tid]; (void)re; (void)im; (void)f;\n"
           "static double RE[QBJ_NT][QBJ_NR], IM[QBJ_NT][QBJ_NR], SM[2 << QBJ_T];\n"
           "static u32 F[QBJ_NT];\n"
           "static u16 sidx_tab["
        << ntab
        << " * QBJ_NT];\nstatic u64 goff_tab[2 * QBJ_NT];\nstatic u32 line_tab[(1 << (QBJ_T - 3))];\nstatic u16 lin_tab[QBJ_NT];\n"
           "extern \"C\" int qb_jit_pass_host(double *dst, const double *amps, u64 ntiles, const QbjArgs *Ap, u64 args_bytes) {\n"
           "  if (args_bytes != sizeof(QbjArgs)) return -1;\n"
           "  const QbjArgs &A = *Ap;\n"
           "  for (u32 tid = 0; tid < QBJ_NT; ++tid) {\n";
      emit_tables();
      o << "  }\n  (void)line_tab;\n  (void)lin_tab;\n"
           "  for (u64 tile_id = 0; tile_id < ntiles; ++tile_id) {\n"
           "  u64 base;\n";
      deposit("base", "tile_id");
      o << "  const u64 basefull = base | A.rank_bits;\n  (void)basefull;\n";
      if (tma)  // what the copy engine does: line c of the tile lands at byte 128 c of the buffer
        o << "  for (u32 c_ = 0; c_ < (1u << (QBJ_T - 3)); ++c_)\n"
             "    for (u32 e_ = 0; e_ < 16; ++e_) SM[16 * c_ + e_] = amps[2 * (base + ((u64)line_tab[c_] << 3)) + e_];\n";
      o << "  for (u32 tid = 0; tid < QBJ_NT; ++tid) {\n    QBJ_THREAD_REFS\n    f = 0;\n";
      if (tma) emit_load_smem(); else emit_load();
      o << rounds_txt;
      if (oop) emit_obase(o, "tile_id", "    ");
      emit_store(mf_end);
      o << "  }\n  }\n  return 0;\n}\n";
    }
  }
  if (bad) return fail("step structure the generator does not accept");
  out.key.swap(g.key);
  out.coefs.swap(g.coefs);
  out.left_out = g.left_out;
  out.src = g.want_src ? g.o.str() : std::string();
  out.T = T;
  out.R = R;
  out.minb = minb;
  out.nrounds = nrounds;
  out.swz_fixed = swz_fixed;
  out.swz_conflicts = swz_conflicts;
  out.smem = smem;
  out.threads = dual ? 2 * NT : NT;
  out.args_bytes = sizeof(JitArgsHead) + sizeof(double) * std::max<size_t>(1, out.coefs.size());
  if (!std::isfinite(out.left_out) || out.left_out == 0.0) return fail("deferred factor out of range");
  return true;
}

bool jit_quick(const PassPlan &pp, JitProgram &out, std::string *why) {
  auto fail = [&](const char *m) {
    if (why) *why = m;
    return false;
  };
  // ---- the same eligibility rules as jit_generate
  if (pp.blob.size() < sizeof(DevPass)) return fail("short blob");
  const DevPass &P = *reinterpret_cast<const DevPass *>(pp.blob.data());
  if (P.lite == 0) return fail("not a step (lite) pass");
  if (P.dbg_skip && !(P.dbg_skip & 16u)) return fail("profiling switches set");
  if ((P.dbg_skip & 15u) && (P.tma != 0 || P.jit_group > 1)) return fail("profiling switches: plain tile loop only");
  if (pp.blob.size() < sizeof(DevPass) + size_t(P.nsteps) * sizeof(DevStep)) return fail("short blob");
  const DevStep *S = reinterpret_cast<const DevStep *>(pp.blob.data() + sizeof(DevPass));
  const int T = (int)P.tile_bits, R = (int)P.reg_bits;
  const int nrounds = (int)P.nrounds;
  if (R < 3 || R > kMaxRegBits || T > kMaxTileBits || T - R > 10 || nrounds < 1 || nrounds > kMaxRounds)
    return fail("unsupported geometry");
  if (!(P.tile_pos[0] == 0 && P.tile_pos[1] == 1 && P.tile_pos[2] == 2)) return fail("tile without the low line bits");
  // ---- digest: two independent 64-bit walks over every field the generator reads
  uint64_t h1 = 1469598103934665603ull, h2 = 0x9E3779B97F4A7C15ull;
  auto mix = [&](uint64_t v) {
    h1 = (h1 ^ v) * 1099511628211ull;
    h1 ^= h1 >> 32;
    h2 = (h2 + v) * 0xFF51AFD7ED558CCDull;
    h2 ^= h2 >> 29;
  };
  mix(0x716a6231);  // format tag
  mix(P.tile_bits); mix(P.reg_bits); mix(P.nrounds); mix(P.l2_prefetch); mix(P.has_gscale != 0);
  mix(P.nruns);
  for (uint32_t k = 0; k < P.nruns && k < (uint32_t)kMaxRuns; ++k) { mix(P.run_shift[k]); mix(P.run_len[k]); }
  for (int i = 0; i < T; ++i) mix(P.tile_pos[i]);
  mix(P.oop);
  mix(P.oop != 0 && P.xch.n != 0);
  for (int i = 0; i < T; ++i) mix(P.out_pos[i]);
  mix(P.onruns);
  for (uint32_t k = 0; k < P.onruns && k < (uint32_t)kMaxOutRuns; ++k) { mix(P.orun_len[k]); mix(P.orun_shift[k]); }
  mix(P.jit_group); mix(P.jit_pf_last); mix(P.jit_minb); mix(P.jit_mem); mix(P.nsteps); mix(P.tma); mix(P.dbg_skip); mix(P.pf_lines);
  out.coefs.clear();
  double left = 1.0;
  bool bad = false;
  uint32_t mf = 0;
  for (int r = 0; r < nrounds; ++r) {
    const DevRound &RD = P.rounds[r];
    for (int j = 0; j < T - R; ++j) mix(RD.tid_pos[j]);
    for (int j = 0; j < R; ++j) mix(RD.reg_pos[j]);
    mix(RD.warp_local); mix(RD.step_begin); mix(RD.step_end);
    if (r > 0) mf = 0;  // flips are folded into the transpose
    if (RD.step_end > P.nsteps || RD.step_begin > RD.step_end) return fail("step range");
    for (uint32_t si = RD.step_begin; si < RD.step_end; ++si) {
      const DevStep &st = S[si];
      mix(st.kinds); mix(st.ntog); mix(st.swap_j);
      for (int J = 0; J < R; ++J) {
        const uint32_t kind = (st.kinds >> (4 * J)) & 15u;
        const uint32_t cls = kind & SLOT_CLASS;
        if (cls == SLOT_NONE) continue;
        const bool flip = ((mf >> J) & 1u) != 0;
        if (flip && !(kind & SLOT_FLIP)) bad = true;
        if (cls == SLOT_ROT) {
          const double cs = st.slot[J][2], sn = st.slot[J][3];
          if (!(cs >= 0.0) || !std::isfinite(cs) || !std::isfinite(sn) || !(cs * cs + sn * sn > 0.5)) bad = true;
          const bool formA = std::fabs(sn) <= cs;
          mix(formA);
          if (formA) {
            out.coefs.push_back(sn / cs);
            left *= cs;
          } else if (!flip) {
            out.coefs.push_back(cs / sn);
            left *= sn;
          } else {
            out.coefs.push_back(st.slot[J][0]);
            out.coefs.push_back(st.slot[J][1]);
          }
        } else if (cls == SLOT_REAL) {
          for (int e = 0; e < 4; ++e) out.coefs.push_back(st.slot[J][e]);
        } else if (cls == SLOT_GENERAL || cls == SLOT_GENERAL1) {
          for (int e = 0; e < 8; ++e) out.coefs.push_back(st.slot[J][e]);
        } else {
          bad = true;
        }
      }
      if (st.ntog > (uint32_t)kStepToggles) bad = true;
      for (uint32_t k = 0; k < st.ntog && k < (uint32_t)kStepToggles; ++k) {
        const auto &tg = st.tog[k];
        mix(tg.cthr); mix(tg.bit); mix(tg.cext);
        if ((int)tg.bit >= R) bad = true;
        if (tg.cthr == 0 && tg.cext == 0) continue;  // uncontrolled X: a renaming, no flip
        mf |= 1u << tg.bit;
      }
      if (st.swap_j != 0xffu) {
        const int J = (int)(st.swap_j & 7u);
        mix(st.swap_creg); mix(st.swap_cthr); mix(st.swap_cext);
        if (J >= R || st.swap_creg == 0 || (st.swap_creg >> R) != 0 || ((st.swap_creg >> J) & 1u)) bad = true;
      }
    }
  }
  if (bad) return fail("step structure the generator does not accept");
  if (!std::isfinite(left) || left == 0.0) return fail("deferred factor out of range");
  const uint64_t nco = out.coefs.size();
  out.key.assign(24, '\0');
  std::memcpy(&out.key[0], &h1, 8);
  std::memcpy(&out.key[8], &h2, 8);
  std::memcpy(&out.key[16], &nco, 8);
  out.left_out = left;
  out.src.clear();
  out.T = T;
  out.R = R;
  out.nrounds = nrounds;
  out.args_bytes = sizeof(JitArgsHead) + sizeof(double) * std::max<size_t>(1, out.coefs.size());
  return true;
}

std::vector<uint8_t> jit_pack_args(const JitProgram &p, const double gs[2], uint64_t rank_bits, uint64_t base_fixed,
                                   const XchGeom *xch) {
  std::vector<uint8_t> a(p.args_bytes, 0);
  JitArgsHead h;
  std::memset(&h, 0, sizeof h);
  if (xch) h.xch = *xch;
  h.gs[0] = gs[0];
  h.gs[1] = gs[1];
  h.rank_bits = rank_bits;
  h.base_fixed = base_fixed;
  std::memcpy(a.data(), &h, sizeof(h));
  if (!p.coefs.empty()) std::memcpy(a.data() + sizeof(h), p.coefs.data(), sizeof(double) * p.coefs.size());
  return a;
}

}  // namespace qb

// =================================================================== runtime (libqubism_sv.so only)
#ifndef QB_JIT_NO_RUNTIME
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace qb {
namespace {

// ---- NVRTC, resolved at run time: the library must load (and the generic kernels run) on a box
// without it
typedef struct _nvrtcProgram *nvrtcProgram_t;
struct Nvrtc {
  void *lib = nullptr;
  int (*CreateProgram)(nvrtcProgram_t *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*CompileProgram)(nvrtcProgram_t, int, const char *const *) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram_t, size_t *) = nullptr;
  int (*GetCUBIN)(nvrtcProgram_t, char *) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram_t, size_t *) = nullptr;
  int (*GetProgramLog)(nvrtcProgram_t, char *) = nullptr;
  int (*DestroyProgram)(nvrtcProgram_t *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*Version)(int *, int *) = nullptr;
  bool wide_ldst = false;  // ld / st .v4.f64 (256 bits): PTX ISA 8.8 = CUDA 12.9
  std::string why;
  bool ok = false;
};

Nvrtc &nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    // the toolkit's own NVRTC first: a Python process that imported torch already holds torch's
    // bundled libnvrtc.so.12 (CUDA 12.8), whose ptxas rejects the 256-bit ld / st of sm_100a
    const char *names[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12",
                           "libnvrtc.so"};
    for (const char *nm : names) {
      n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.lib) break;
    }
    if (!n.lib) {
      n.why = "libnvrtc not found";
      return;
    }
#define QB_SYM(field, name)                                   \
  *reinterpret_cast<void **>(&n.field) = dlsym(n.lib, name);  \
  if (!n.field) {                                             \
    n.why = std::string("libnvrtc lacks ") + name;            \
    return;                                                   \
  }
    QB_SYM(CreateProgram, "nvrtcCreateProgram")
    QB_SYM(CompileProgram, "nvrtcCompileProgram")
    QB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    QB_SYM(GetCUBIN, "nvrtcGetCUBIN")
    QB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    QB_SYM(GetProgramLog, "nvrtcGetProgramLog")
    QB_SYM(DestroyProgram, "nvrtcDestroyProgram")
    QB_SYM(GetErrorString, "nvrtcGetErrorString")
    QB_SYM(Version, "nvrtcVersion")
#undef QB_SYM
    int major = 0, minor = 0;
    if (n.Version(&major, &minor) == 0) n.wide_ldst = major > 12 || (major == 12 && minor >= 9);
    n.ok = true;
  });
  return n;
}

// ---- the driver entry points, through the runtime (no link against libcuda)
struct Driver {
  CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                           void **) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
  std::string why;
  bool ok = false;
};

Driver &driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [&](const char *name, void **fn) {
      cudaDriverEntryPointQueryResult q;
      const cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
      if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
        d.why = std::string("driver entry point ") + name + " unavailable";
        (void)cudaGetLastError();
        return false;
      }
      return true;
    };
    if (!get("cuModuleLoadData", reinterpret_cast<void **>(&d.ModuleLoadData))) return;
    if (!get("cuModuleGetFunction", reinterpret_cast<void **>(&d.ModuleGetFunction))) return;
    if (!get("cuFuncSetAttribute", reinterpret_cast<void **>(&d.FuncSetAttribute))) return;
    if (!get("cuOccupancyMaxActiveBlocksPerMultiprocessor",
             reinterpret_cast<void **>(&d.OccupancyMaxActiveBlocksPerMultiprocessor)))
      return;
    if (!get("cuLaunchKernel", reinterpret_cast<void **>(&d.LaunchKernel))) return;
    if (!get("cuGetErrorString", reinterpret_cast<void **>(&d.GetErrorString))) return;
    d.ok = true;
  });
  return d;
}

std::string cu_err(CUresult r) {
  const char *s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  return s ? std::string(s) : ("CUresult " + std::to_string((int)r));
}

struct Entry {
  int seen = 0;
  int state = 0;  // 0 not compiled, 1 ready, -1 failed (stay on the generic kernel), 2 compiling, 3 cubin ready
  std::vector<char> cubin;
  std::string err;
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  int occ = 0, threads = 0;
  size_t smem = 0;
};
std::mutex g_mu;
std::unordered_map<std::string, Entry> g_cache;
JitStats g_stats;

bool compile_cubin(const std::string &src, std::vector<char> &cubin, std::string *err) {
  Nvrtc &n = nvrtc();
  nvrtcProgram_t prog = nullptr;
  int rc = n.CreateProgram(&prog, src.c_str(), "qb_jit_pass.cu", 0, nullptr, nullptr);
  if (rc != 0) {
    if (err) *err = std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc);
    return false;
  }
  const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-DQBJ_NO_LD256=1"};
  rc = n.CompileProgram(prog, n.wide_ldst ? 3 : 4, opts);
  if (rc != 0) {
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) n.GetProgramLog(prog, &log[0]);
    if (err) *err = std::string("nvrtcCompileProgram: ") + n.GetErrorString(rc) + "\n" + log.substr(0, 2000);
    n.DestroyProgram(&prog);
    return false;
  }
  size_t cs = 0;
  rc = n.GetCUBINSize(prog, &cs);
  if (rc == 0 && cs) {
    cubin.resize(cs);
    rc = n.GetCUBIN(prog, cubin.data());
  }
  n.DestroyProgram(&prog);
  if (rc != 0 || cubin.empty()) {
    if (err) *err = "nvrtcGetCUBIN failed";
    return false;
  }
  return true;
}

}  // namespace

std::string jit_toolchain() {
  Nvrtc &n = nvrtc();
  if (!n.ok) return "nvrtc unavailable (" + n.why + "): generic kernels only";
  int major = 0, minor = 0;
  n.Version(&major, &minor);
  return "nvrtc " + std::to_string(major) + "." + std::to_string(minor) +
         (n.wide_ldst ? ", 256-bit ld/st.global.v4.f64" : ", 128-bit ld/st (PTX ISA < 8.8: no .v4.f64)");
}

bool jit_available(std::string *why) {
  if (!nvrtc().ok) {
    if (why) *why = nvrtc().why;
    return false;
  }
  if (!driver().ok) {
    if (why) *why = driver().why;
    return false;
  }
  return true;
}

// test hook: source -> cubin without touching a device
bool jit_compile_only(const std::string &src, size_t *cubin_bytes, std::string *err) {
  if (!nvrtc().ok) {
    if (err) *err = nvrtc().why;
    return false;
  }
  std::vector<char> cubin;
  if (!compile_cubin(src, cubin, err)) return false;
  if (cubin_bytes) *cubin_bytes = cubin.size();
  return true;
}

namespace {

// ---- background compilation: NVRTC runs on worker threads (it is CPU work and thread-safe);
// every CUDA driver call stays on the caller's thread (module load at the next sighting)
struct Job {
  Entry *e;
  PassPlan pp;  // blob only
};
struct Pool {
  std::mutex mu;
  std::condition_variable cv, idle;
  std::deque<Job> jobs;
  int inflight = 0;
  int nworkers = 0;
};
Pool &pool() {
  static Pool *p = new Pool;  // never destroyed: workers may outlive static destruction
  return *p;
}

bool build_cubin(const PassPlan &pp, std::vector<char> &cubin, JitProgram &full, std::string *err) {
  std::string why;
  if (!jit_generate(pp, JIT_DEVICE_SRC, full, &why)) {
    if (err) *err = "generator: " + why;
    return false;
  }
  return compile_cubin(full.src, cubin, err);
}

void worker_main() {
  Pool &P = pool();
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(P.mu);
      P.cv.wait(lk, [&] { return !P.jobs.empty(); });
      job = std::move(P.jobs.front());
      P.jobs.pop_front();
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<char> cubin;
    JitProgram full;
    std::string err;
    const bool ok = build_cubin(job.pp, cubin, full, &err);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    {
      std::lock_guard<std::mutex> lock(g_mu);
      if (ok) {
        job.e->cubin.swap(cubin);
        job.e->smem = full.smem;
        job.e->threads = full.threads;
        job.e->state = 3;
      } else {
        job.e->state = -1;
        job.e->err = err;
        g_stats.failed++;
      }
      g_stats.compile_ms += ms;
    }
    {
      std::lock_guard<std::mutex> lk(P.mu);
      --P.inflight;
    }
    P.idle.notify_all();
  }
}

// g_mu held.  cubin -> module -> function; state 1 or -1
bool load_entry(Entry &e, std::string *err) {
  Driver &d = driver();
  CUresult r = d.ModuleLoadData(&e.mod, e.cubin.data());
  if (r == CUDA_SUCCESS) r = d.ModuleGetFunction(&e.fn, e.mod, "qb_jit_pass");
  if (r == CUDA_SUCCESS) r = d.FuncSetAttribute(e.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)e.smem);
  if (r == CUDA_SUCCESS) r = d.OccupancyMaxActiveBlocksPerMultiprocessor(&e.occ, e.fn, e.threads, e.smem);
  std::vector<char>().swap(e.cubin);
  if (r != CUDA_SUCCESS || e.occ < 1) {
    e.state = -1;
    g_stats.failed++;
    if (err) *err = "loading the specialised kernel: " + (r != CUDA_SUCCESS ? cu_err(r) : std::string("occupancy 0"));
    return false;
  }
  e.state = 1;
  g_stats.compiled++;
  return true;
}

}  // namespace

// The cache is keyed by a 128-bit digest of (device, structural key): a long run over ever new
// structures (an interpreter session, a sharded state whose layout never comes back) must not
// keep ~10 KB of key per pass it has ever planned.
static std::string digest_of(int dev, const std::string &key) {
  uint64_t h1 = 1469598103934665603ull ^ (uint64_t)dev, h2 = 0x9E3779B97F4A7C15ull + (uint64_t)dev;
  for (unsigned char ch : key) {
    h1 = (h1 ^ ch) * 1099511628211ull;                               // FNV-1a
    h2 = (h2 + ch) * 0xFF51AFD7ED558CCDull;                          // an independent multiply-xorshift walk
    h2 ^= h2 >> 29;
  }
  const uint64_t len = key.size();
  std::string d(24, '\0');
  std::memcpy(&d[0], &h1, 8);
  std::memcpy(&d[8], &h2, 8);
  std::memcpy(&d[16], &len, 8);
  return d;
}
constexpr size_t kMaxSightings = 1u << 16;  // sighting records kept before the never-compiled ones are dropped
constexpr uint64_t kMaxCompiled = 4096;     // specialised kernels per process

// Sighting counts are kept per (device, salt, structure): the salt is the rank of an in-process rank
// group, where P ranks share this cache -- each of them must see the sequence of counts a rank
// process of its own would see (`requested` has to come out the same on every rank).  Compiled
// kernels are shared by all ranks of the device.
struct Sighting {
  int seen = 0;
  bool requested = false;
};
static std::unordered_map<std::string, Sighting> g_seen;

int jit_lookup(const JitProgram &kp, const PassPlan &pp, int threshold, void **handle, std::string *err, bool *requested,
               int salt) {
  int dev = 0;
  if (requested) *requested = false;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  const std::string key = digest_of(dev, kp.key);
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_cache.size() >= kMaxSightings && g_cache.find(key) == g_cache.end()) {
    for (auto it = g_cache.begin(); it != g_cache.end();)
      it = (it->second.state == 0) ? g_cache.erase(it) : std::next(it);  // (entries being compiled / loaded stay: workers hold pointers)
  }
  if (g_seen.size() >= kMaxSightings) {
    for (auto it = g_seen.begin(); it != g_seen.end();) it = it->second.requested ? std::next(it) : g_seen.erase(it);
  }
  Entry &e = g_cache[key];
  std::string skey = key;
  skey.append(reinterpret_cast<const char *>(&salt), sizeof salt);
  Sighting &sg = g_seen[skey];
  // `requested` depends only on how often THIS rank looked this structure up: every rank of a
  // sharded state sees the same sequence of structures, so it is the same on all of them -- unlike
  // the moment a background compilation finishes
  if (!sg.requested && ++sg.seen >= threshold) sg.requested = true;
  if (requested) *requested = sg.requested;
  if (e.state == 3 && !load_entry(e, err)) return -1;
  if (!sg.requested) return 0;
  if (e.state == 1) {
    *handle = &e;
    return 1;
  }
  if (e.state != 0) return 0;  // failed before, or still compiling: the generic kernel runs
  if (g_stats.compiled + g_stats.failed >= kMaxCompiled) return 0;
  if (!jit_available(err)) {
    e.state = -1;
    g_stats.failed++;
    return -1;
  }
  if (threshold <= 1) {  // "specialise at first sight": compile here and now
    const auto t0 = std::chrono::steady_clock::now();
    JitProgram full;
    if (!build_cubin(pp, e.cubin, full, err)) {
      e.state = -1;
      g_stats.failed++;
      return -1;
    }
    e.smem = full.smem;
    e.threads = full.threads;
    const bool ok = load_entry(e, err);
    g_stats.compile_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!ok) return -1;
    *handle = &e;
    return 1;
  }
  // a structure that came back: compile it in the background, keep running the generic kernel
  e.state = 2;
  Pool &P = pool();
  {
    // Process exit must not tear libnvrtc down under a worker that is still compiling.  This
    // handler is registered after libnvrtc was loaded (jit_available above), so exit() runs it
    // BEFORE libnvrtc's own destructors: it simply waits for the queue to drain (bounded).
    static std::once_flag at_exit_once;
    std::call_once(at_exit_once, [] {
      std::atexit([] {
        Pool &Q = pool();
        std::unique_lock<std::mutex> lk(Q.mu);
        Q.idle.wait_for(lk, std::chrono::seconds(30), [&] { return Q.inflight == 0; });
      });
    });
  }
  {
    std::lock_guard<std::mutex> lk(P.mu);
    Job j;
    j.e = &e;
    j.pp.blob = pp.blob;
    P.jobs.push_back(std::move(j));
    ++P.inflight;
    unsigned hw = std::thread::hardware_concurrency();
    const int want = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 2u));
    while (P.nworkers < want && P.nworkers < (int)P.jobs.size()) {
      std::thread(worker_main).detach();
      ++P.nworkers;
    }
  }
  P.cv.notify_one();
  return 0;
}

void jit_wait() {
  Pool &P = pool();
  std::unique_lock<std::mutex> lk(P.mu);
  P.idle.wait(lk, [&] { return P.inflight == 0; });
}

int jit_launch(void *handle, void *amps, const void *src, uint64_t ntiles, const std::vector<uint8_t> &args, int sm_count,
               void *stream, std::string *err) {
  Entry &e = *static_cast<Entry *>(handle);
  uint64_t grid = uint64_t(sm_count) * uint64_t(e.occ);
  if (grid > ntiles) grid = ntiles;
  if (grid == 0) return 0;
  unsigned long long nt = ntiles;
  if (!src) src = amps;
  void *params[] = {&amps, &src, &nt, const_cast<uint8_t *>(args.data())};
  const CUresult r = driver().LaunchKernel(e.fn, (unsigned)grid, 1, 1, (unsigned)e.threads, 1, 1, (unsigned)e.smem,
                                           static_cast<CUstream>(stream), params, nullptr);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "launching the specialised kernel: " + cu_err(r);
    return -1;
  }
  {
    std::lock_guard<std::mutex> lock(g_mu);
    g_stats.launches++;
  }
  return 0;
}

JitStats jit_stats() {
  std::lock_guard<std::mutex> lock(g_mu);
  return g_stats;
}

}  // namespace qb
#endif  // QB_JIT_NO_RUNTIME
