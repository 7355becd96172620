// Host emulation of the register-radix transform kernels (snac_b200/csrc/fft_radix.cuh): the
// __host__ __device__ phase functions are run thread by thread, phase by phase (a phase boundary =
// the kernel's __syncthreads), against the dense definitions of the FFTW r2r kinds in fft.cu
// (reference call sites: src/fft.f90:85-149).  TEST INFRASTRUCTURE: checks the index algebra, the
// twiddles and the split-array addressing without a GPU; the GPU parity tests check the kernels.
#include <cmath>
#include <cstdio>
#include <random>

#include "../../snac_b200/csrc/fft.cu"

using namespace snac;

template <int R1, int R2>
static void emulate(const R2RArgs &a, int dir, long ntiles) {
  using P = radix::Plan<R1, R2>;
  std::vector<double2> Z((size_t)P::M * P::TP);
  for (long tile = 0; tile < ntiles; ++tile) {
    for (auto &z : Z) z = make_double2(NAN, NAN);
    auto phase = [&](auto fn) {
      for (int tid = 0; tid < P::THREADS; ++tid) fn(radix::thread_of<P::TP>(a, tid, P::THREADS, (int)tile));
    };
    const bool stage = dir > 0 ? a.in_lane_t != 0 : a.out_lane_t != 0;   // the transform axis is at unit stride on the physical side
    auto stage_phase = [&](bool in) {
      for (int tid = 0; tid < P::THREADS; ++tid) radix::stage_io<R1, R2>(a, Z.data(), tid, P::THREADS, (int)tile, in);
    };
    if (dir > 0) {
      if (stage) {
        stage_phase(true);
        phase([&](const radix::Thread &T) { radix::fwd_pass1<R1, R2, true, true>(a, Z.data(), T); });
      } else
      phase([&](const radix::Thread &T) { radix::fwd_pass1<R1, R2, true>(a, Z.data(), T); });
      phase([&](const radix::Thread &T) { radix::run_pass<R1, R2, false>(a, Z.data(), T); });
      phase([&](const radix::Thread &T) { radix::fwd_untangle<R1, R2, true>(a, Z.data(), T); });
    } else {
      phase([&](const radix::Thread &T) { radix::bwd_tangle<R1, R2, true>(a, Z.data(), T); });
      phase([&](const radix::Thread &T) { radix::run_pass<R1, R2, true>(a, Z.data(), T); });
      if (stage) {
        phase([&](const radix::Thread &T) { radix::bwd_pass2<R1, R2, true, true>(a, Z.data(), T); });
        stage_phase(false);
      } else
      phase([&](const radix::Thread &T) { radix::bwd_pass2<R1, R2, true>(a, Z.data(), T); });
    }
  }
}

static bool run(const R2RArgs &a, int dir, long ntiles) {
  switch (a.nt) {
    case 32: emulate<4, 4>(a, dir, ntiles); return true;
    case 64: emulate<8, 4>(a, dir, ntiles); return true;
    case 128: emulate<8, 8>(a, dir, ntiles); return true;
    case 256: emulate<16, 8>(a, dir, ntiles); return true;
    case 512: emulate<16, 16>(a, dir, ntiles); return true;
  }
  return false;
}

int main() {
  std::mt19937_64 rng(1234);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  const int kinds[] = {K_R2HC, K_HC2R, K_REDFT10, K_REDFT01, K_RODFT10, K_RODFT01};
  double worst = 0.0;
  int cases = 0;
  for (int n : {32, 64, 128, 256, 512})
    for (int axis : {0, 1, 2})   // axis 0: unit stride on the physical side, the tile is staged through shared memory
      for (int kind : kinds)
        for (int split : {1, 2, -2, -4}) {   // one array; two consecutive chunks; cyclic over two / four arrays
          // box: 21 pencils along pd (a partial second tile), 3 along the other direction
          int n3[3] = {3, 3, 3};
          const int pd = axis == 0 ? 1 : 0, od = 3 - axis - pd;
          n3[pd] = 21;
          n3[axis] = n;
          const bool fwd = (kind == K_R2HC || kind == K_REDFT10 || kind == K_RODFT10);
          // physical array: ghosted, padded; spectral array: (x, other, q) as the solver wants it
          const long psy = n3[0] + 19, psz = psy * (n3[1] + 2);
          const long pst[3] = {1, psy, psz};
          long sst[3];
          sst[pd] = 1; sst[od] = 24; sst[axis] = 24L * n3[od];
          const size_t pe = (size_t)psz * (n3[2] + 2) + 64, se = (size_t)sst[axis] * n + 64;
          std::vector<double> phys(pe, 0.0), spec(se, 0.0), spec2(se, 0.0), spec3(se, 0.0), spec4(se, 0.0);
          std::vector<double> *pieces[4] = {&spec, &spec2, &spec3, &spec4};
          const int ncyc = split < 0 ? -split : 0;
          const long pint = 1 + psy + psz;
          const std::vector<double> tab = r2r_fast_tables(n);
          const std::vector<double> Mx = dense_matrix(kind, n);
          R2RArgs a;
          for (int p = 0; p < 16; ++p) { a.inp[p] = nullptr; a.outp[p] = nullptr; }
          const int chunk = split == 2 ? n / 2 : n;
          // the split side is the spectral one: two "owners", each holding half of the wavenumbers
          // in its own array (here: the same layout, second half in spec2)
          for (int d = 0; d < 3; ++d) { a.n[d] = n3[d]; a.ist[d] = fwd ? pst[d] : sst[d]; a.ost[d] = fwd ? sst[d] : pst[d]; }
          a.axis = axis; a.nt = n; a.tp = radix::tp_of(n); a.pd = pd; a.od = od; a.kind = kind;
          a.in_lane_t = (fwd && axis == 0) ? 1 : 0; a.out_lane_t = (!fwd && axis == 0) ? 1 : 0; a.scale = 0.75; a.tab = tab.data(); a.chunk = chunk;
          a.ns_mask = ncyc ? ncyc - 1 : 0; a.ns_shift = ncyc == 4 ? 2 : (ncyc == 2 ? 1 : 0);
          std::vector<double> in((size_t)n3[0] * n3[1] * n3[2]);
          for (auto &v : in) v = U(rng);
          auto lin = [&](int i, int j, int k) { return ((size_t)k * n3[1] + j) * n3[0] + i; };
          auto spec_at = [&](int i, int j, int k) -> double & {
            const int idx[3] = {i, j, k};
            const int t = idx[axis];
            if (ncyc) {   // index t lives in piece t mod ncyc at position t div ncyc
              int id2[3] = {i, j, k};
              id2[axis] = t / ncyc;
              return (*pieces[t % ncyc])[id2[0] * sst[0] + id2[1] * sst[1] + id2[2] * sst[2]];
            }
            std::vector<double> &arr = (split > 1 && t >= chunk) ? spec2 : spec;
            return arr[idx[0] * sst[0] + idx[1] * sst[1] + idx[2] * sst[2]];
          };
          if (fwd) {
            for (int k = 0; k < n3[2]; ++k) for (int j = 0; j < n3[1]; ++j) for (int i = 0; i < n3[0]; ++i)
              phys[pint + i + psy * j + psz * k] = in[lin(i, j, k)];
            a.inp[0] = phys.data() + pint; a.in_split = 0;
            a.outp[0] = spec.data(); a.outp[1] = spec2.data(); a.outp[2] = spec3.data(); a.outp[3] = spec4.data();
            a.out_split = ncyc ? 2 : (split > 1 ? 1 : 0);
          } else {
            for (int k = 0; k < n3[2]; ++k) for (int j = 0; j < n3[1]; ++j) for (int i = 0; i < n3[0]; ++i)
              spec_at(i, j, k) = in[lin(i, j, k)];
            a.inp[0] = spec.data(); a.inp[1] = spec2.data(); a.inp[2] = spec3.data(); a.inp[3] = spec4.data();
            a.in_split = ncyc ? 2 : (split > 1 ? 1 : 0);
            a.outp[0] = phys.data() + pint; a.out_split = 0;
          }
          const long ntiles = (long)((n3[pd] + radix::tp_of(n) - 1) / radix::tp_of(n)) * n3[od];
          if (!run(a, fwd ? +1 : -1, ntiles)) { printf("no kernel for n = %d\n", n); return 1; }
          // reference: dense matrix-vector product along the axis
          double err = 0.0, nrm = 0.0;
          for (int k = 0; k < n3[2]; ++k) for (int j = 0; j < n3[1]; ++j) for (int i = 0; i < n3[0]; ++i) {
            int idx[3] = {i, j, k};
            const int t = idx[axis];
            double acc = 0.0;
            for (int s = 0; s < n; ++s) {
              int id2[3] = {i, j, k};
              id2[axis] = s;
              acc += Mx[(size_t)t * n + s] * in[lin(id2[0], id2[1], id2[2])];
            }
            acc *= a.scale;
            const double got = fwd ? spec_at(i, j, k) : phys[pint + i + psy * j + psz * k];
            err = std::max(err, std::fabs(got - acc));
            nrm = std::max(nrm, std::fabs(acc));
          }
          const double rel = err / nrm;
          worst = std::max(worst, rel);
          ++cases;
          if (!(rel < 1e-13 * std::log2((double)n))) {
            printf("FAIL n=%d axis=%d kind=%d split=%d rel=%.3e\n", n, axis, kind, split, rel);
            return 1;
          }
        }
  printf("OK %d cases, worst relative error %.3e\n", cases, worst);
  return 0;
}
