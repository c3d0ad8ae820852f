// Structure-specialised fused passes (see qb_jit_prelude.cuh for the why).
//
// jit_generate() turns a LITE PassPlan (steps of 1-qubit slots, toggles and register-controlled X:
// everything the reference interpreter ever sends, QASM/Simulation.hs:94-122,163-171) into
//   * a structural KEY (every literal that shapes the code; no gate coefficient),
//   * the run-time coefficient vector in the order the code reads it,
//   * the real factor the kernel leaves OUT of the amplitudes (product of the rotations'
//     deferred cosines / sines): the caller multiplies it into the state's deferred scalar,
//   * on demand the CUDA source (device) or a C++ emulation of it (host, tests only).
// The runtime half (NVRTC through dlopen, module cache, launch) only exists in libqubism_sv.so.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "qb_internal.h"

namespace qb {

struct JitProgram {
  std::string key;             // structural key (binary); equal keys <=> identical source
  std::vector<double> coefs;   // QbjArgs::c
  double left_out = 1.0;       // true amplitudes = left_out * what the kernel writes
  std::string src;             // only if requested
  int T = 0, R = 0, minb = 2, nrounds = 0;
  int swz_fixed = 0;           // transposes whose swizzle columns were re-chosen to remove a bank conflict
  int swz_conflicts = 0;       // transposes left with a 2-way conflict (no assignment found)
  size_t smem = 0;             // dynamic shared memory of the device kernel
  int threads = 0;             // threads per CTA of the device kernel
  size_t args_bytes = 0;       // sizeof(QbjArgs)
};
enum JitEmit { JIT_KEY_ONLY = 0, JIT_DEVICE_SRC = 1, JIT_HOST_SRC = 2 };

// false: this pass cannot be specialised (not a lite pass, inconsistent flip tracking, ...).
bool jit_generate(const PassPlan &pp, JitEmit emit, JitProgram &out, std::string *why);
// What a flush needs per pass and per launch, without building a single string: a 128-bit digest of
// the structure (the same inputs jit_generate reads), the coefficient vector in the generator's
// order and the factor left out.  ~2 us per pass where the full key walk costs ~80 (a 20-qubit
// pass runs in 10 us).  tests/emul cross-checks it against jit_generate on every pass it plans.
bool jit_quick(const PassPlan &pp, JitProgram &out, std::string *why);
// the prelude text (qb_jit_prelude.cuh, embedded at build time)
const char *jit_prelude();

// layout of the kernel parameter block the generated code declares (c[] has coefs.size() entries, >= 1)
struct JitArgsHead {
  double gs[2];
  uint64_t rank_bits;
  uint64_t base_fixed;
  XchGeom xch;  // run-time geometry of a swap fused into the stores (peer pointers differ from process to process)
};
std::vector<uint8_t> jit_pack_args(const JitProgram &p, const double gs[2], uint64_t rank_bits, uint64_t base_fixed,
                                   const XchGeom *xch = nullptr);

#ifndef QB_JIT_NO_RUNTIME
struct JitStats {
  uint64_t compiled = 0, launches = 0, failed = 0;
  double compile_ms = 0.0;
};
// Looks the structure up.  threshold <= 1: compile at first sight, synchronously.  Otherwise a
// structure seen `threshold` times is handed to background threads (NVRTC is CPU work) and the
// generic kernel keeps running until the cubin is there.  Returns 1 when a kernel is ready
// (*handle set), 0 when the caller should use the generic kernel, < 0 on a compile / load error
// (message in *err).
// *requested (may be null): this structure has reached the threshold (now or earlier) -- a
// function of the lookup sequence only, hence identical on every rank of a sharded state.
// salt: the rank inside an in-process rank group (0 otherwise): sightings are counted per rank.
int jit_lookup(const JitProgram &key_only, const PassPlan &pp, int threshold, void **handle, std::string *err,
               bool *requested = nullptr, int salt = 0);
// launch on `stream`; grid = SMs x resident CTAs (capped at ntiles)
// src: where the tiles are read (null / == amps: in place)
int jit_launch(void *handle, void *amps, const void *src, uint64_t ntiles, const std::vector<uint8_t> &args, int sm_count,
               void *stream, std::string *err);
bool jit_available(std::string *why);
std::string jit_toolchain();  // which NVRTC serves the specialised kernels, and the global access width it allows
void jit_wait();  // block until no background compilation is pending
// test hook: source -> cubin with NVRTC, no device needed
bool jit_compile_only(const std::string &src, size_t *cubin_bytes, std::string *err);
JitStats jit_stats();
#endif

}  // namespace qb
