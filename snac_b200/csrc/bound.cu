// K2: halo exchange and boundary-condition face fills as packed device kernels.
//
// Replaces updthalo (reference src/bound.f90:587-650: two blocking MPI_SENDRECV per direction and
// field with derived vector types, src/initmpi.f90:255-271) and set_bc (src/bound.f90:108-248).
// All boxes and all fields of one bounduvw/boundp call are handled by ONE launch per direction:
// a table of plane-copy descriptors (halos between boxes resident on this GPU, including periodic
// self-neighbours) and a table of face-fill descriptors live in device memory, built at commit.
// Directions are processed x, y, z with full ghost extents so that edges and corners propagate
// exactly as in the reference (src/bound.f90:26-34).
#include <cstdlib>

#include "common.h"

namespace snac {

// remote: some destinations are in a peer's memory.  The flag that tells the peer "the push has arrived" is written by
// a LATER kernel (comm.cu); a fence in that kernel does not order THESE threads' stores.  Every writer therefore ends
// with a system-scope fence: its stores have been performed at the peer before the kernel can complete, so the flag
// cannot overtake the data on the way through the NVLink fabric (seen as a run-to-run varying 1e-3 error with
// 32^3 boxes on 4 GPUs, where the peer's unpack kernel reads the landing buffer within microseconds of the flag).
__global__ void copy_ops_kernel(const CopyOp *ops, int remote) {
  const CopyOp op = ops[blockIdx.y];
  const long tot = (long)op.na * op.nb;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
    const int ia = (int)(t % op.na), ib = (int)(t / op.na);
    op.dst[ia * op.dsa + ib * op.dsb] = op.src[ia * op.ssa + ib * op.ssb];
  }
  if (remote) __threadfence_system();
}

__global__ void face_ops_kernel(const FaceOp *ops) {
  const FaceOp op = ops[blockIdx.y];
  const long tot = (long)op.na * op.nb;
  const long sd = op.sd;
  const int n = op.n;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
    const int ia = (int)(t % op.na), ib = (int)(t / op.na);
    double *q = op.p + ia * op.sa + ib * op.sb;
    switch (op.kind) {   // src/bound.f90:135-246
      case FACE_CEN_LO: q[0] = op.factor + op.sgn * q[sd]; break;
      case FACE_CEN_HI: q[(n + 1) * sd] = op.factor + op.sgn * q[n * sd]; break;
      case FACE_STAG_D_LO: q[0] = op.factor; break;
      case FACE_STAG_D_HI: q[n * sd] = op.factor; q[(n + 1) * sd] = q[(n - 1) * sd]; break;
      case FACE_STAG_N_LO: q[0] = op.factor + q[sd]; break;
      case FACE_STAG_N_HI: {
        const double v = op.factor + q[(n - 1) * sd];
        q[n * sd] = v;
        q[(n + 1) * sd] = v;
      } break;
      // inflow (src/bound.f90:456-472)
      case FACE_INFLOW_N_LO: q[0] = op.plane[t]; break;
      case FACE_INFLOW_N_HI: q[n * sd] = op.plane[t]; break;
      case FACE_INFLOW_T_LO: q[0] = 2.0 * op.plane[t] - q[sd]; break;
      case FACE_INFLOW_T_HI: q[(n + 1) * sd] = 2.0 * op.plane[t] - q[n * sd]; break;
    }
  }
}

template <class T>
static int max_plane(const std::vector<T> &v) {
  long m = 1;
  for (auto &o : v) m = std::max(m, (long)o.na * o.nb);
  return (int)std::min<long>((m + 255) / 256, 256);
}

struct OpsMeta { int gx; };

void launch_copy_ops(Ctx &ctx, const DevOps<CopyOp> &ops, bool remote) {
  if (!ops.count) return;
  dim3 grd(ops.count ? 64 : 1, ops.count);
  copy_ops_kernel<<<grd, 256, 0, ctx.stream>>>(ops.d, remote ? 1 : 0);
  ctx.launches++;
}

void launch_face_ops(Ctx &ctx, const DevOps<FaceOp> &ops) {
  if (!ops.count) return;
  dim3 grd(64, ops.count);
  face_ops_kernel<<<grd, 256, 0, ctx.stream>>>(ops.d);
  ctx.launches++;
}

void Exchange::build(Ctx &ctx, const std::vector<ExField> &fields, const std::vector<NbTable> *nbs, const std::vector<int> *keys,
                     bool ids_local, const std::vector<Grid3> *rank_grids, bool stage_x) {
  // `fields` holds, for each exchanged array, one entry per local box, grouped: fields[f*nl + lb].
  // Neighbour on this GPU: the ghost plane of box lb on side ib is pulled from the neighbour's
  // first/last interior plane (src/bound.f90:602-607).  Neighbour owned by another process: this
  // box's first/last interior plane is pushed into the neighbour's ghost plane in the peer's arena.
  const int nl = nbs ? (int)nbs->size() : (int)ctx.boxes.size();
  if (fields.size() % nl) throw Error("Exchange::build: field list is not a multiple of the box count");
  std::vector<CopyOp> ops[3], unp[3];
  for (int d = 0; d < 3; ++d) mask[d] = 0;
  key0 = (keys && ctx.comm.nproc > 1) ? (*keys)[0] : -1;   // only exchanges of published arrays synchronise with peers
  for (size_t f = 0; f < fields.size() / nl; ++f) {
    for (int lb = 0; lb < nl; ++lb) {
      const ExField &me = fields[f * nl + lb];
      for (int d = 0; d < 3; ++d) {
        const int a = (d + 1) % 3, b = (d + 2) % 3;
        const int da = std::min(a, b), db = std::max(a, b);  // keep x (or y) fastest in the plane
        for (int ib = 0; ib < 2; ++ib) {
          const int nbr = nbs ? (*nbs)[lb].nb[d][ib] : ctx.boxes[lb]->info.nb[d][ib];
          if (nbr < 0) continue;
          const int ln = ids_local ? nbr : ctx.local_of_rank[nbr];
          CopyOp op;
          op.na = me.g.n[da] + 2;
          op.nb = me.g.n[db] + 2;
          if (ln >= 0) {
            const ExField &src = fields[f * nl + ln];
            if (src.g.n[da] != me.g.n[da] || src.g.n[db] != me.g.n[db])
              throw Error("halo exchange: neighbouring boxes are not face-congruent");
            op.ssa = src.g.st(da); op.ssb = src.g.st(db);
            op.dsa = me.g.st(da);  op.dsb = me.g.st(db);
            const long sidx = (ib == 0) ? src.g.n[d] : 1;          // neighbour's hi / lo interior plane
            const long didx = (ib == 0) ? 0 : me.g.n[d] + 1;        // my lo / hi ghost plane
            op.src = src.p + sidx * src.g.st(d);
            op.dst = me.p + didx * me.g.st(d);
          } else {
            if (!keys || ctx.comm.nproc == 1) throw Error("halo exchange with a box owned by another process needs snac_comm_connect");
            const int proc = ctx.owner[nbr];
            const snac_box &w = ctx.world[nbr];
            const Grid3 rg = rank_grids ? (*rank_grids)[nbr] : Grid3::make(w.hi[0] - w.lo[0] + 1, w.hi[1] - w.lo[1] + 1, w.hi[2] - w.lo[2] + 1);
            if (rg.n[da] != me.g.n[da] || rg.n[db] != me.g.n[db])
              throw Error("halo exchange: neighbouring boxes are not face-congruent");
            const long sidx = (ib == 0) ? 1 : me.g.n[d];            // my lo / hi interior plane
            const long didx = (ib == 0) ? rg.n[d] + 1 : 0;          // the neighbour's hi / lo ghost plane
            op.ssa = me.g.st(da); op.ssb = me.g.st(db);
            op.src = me.p + sidx * me.g.st(d);
            static const bool xstage = !(getenv("SNAC_XFACE_STAGE") && atoi(getenv("SNAC_XFACE_STAGE")) == 0);
            if (d == 0 && xstage && stage_x) {
              // An x-face is one element per row: pushed in place it is (n2+2)(n3+2) separate 8-byte stores
              // over NVLink.  It is gathered into a dense landing buffer of the neighbour instead (coalesced
              // remote stores); the neighbour scatters it into its ghost column once the push has arrived.
              const int fkey = (*keys)[f];
              op.dst = ctx.comm.remote(proc, nbr, AK_XSTAGE + 2 * fkey + (1 - ib));   // its hi / lo side
              op.dsa = 1; op.dsb = op.na;
              CopyOp un;   // my own ghost plane on side ib is filled from my landing buffer of that side
              un.na = op.na; un.nb = op.nb;
              un.src = ctx.comm.remote(ctx.comm.me, ctx.boxes[lb]->info.rank, AK_XSTAGE + 2 * fkey + ib);
              un.ssa = 1; un.ssb = un.na;
              un.dst = me.p + ((ib == 0) ? 0 : me.g.n[d] + 1) * me.g.st(d);
              un.dsa = me.g.st(da); un.dsb = me.g.st(db);
              unp[d].push_back(un);
            } else {
              double *rp = ctx.comm.remote(proc, nbr, (*keys)[f]) + 15;
              op.dsa = rg.st(da);   op.dsb = rg.st(db);
              op.dst = rp + didx * rg.st(d);
            }
            mask[d] |= 1u << proc;
          }
          ops[d].push_back(op);
        }
      }
    }
  }
  for (int d = 0; d < 3; ++d) { local[d].upload(ops[d]); unpack[d].upload(unp[d]); }
}

// one direction: wait until the neighbours are done with the ghost cells (they reached the same
// exchange), copy / push, then wait for the neighbours' pushes
void Exchange::run(Ctx &ctx, int dir) const {
  if (key0 >= 0) ctx.comm.handshake(ctx, CH_READY, mask[dir]);
  launch_copy_ops(ctx, local[dir], key0 >= 0 && mask[dir] != 0);
  if (key0 >= 0) ctx.comm.handshake(ctx, CH_ARRIVED, mask[dir]);
  launch_copy_ops(ctx, unpack[dir]);
}

// x, then y, then z (edges/corners propagate).  Every direction has its own "ready" handshake: a neighbour may push
// direction d into this box only after this box has finished directions < d.  The x-faces of a remote neighbour land
// in a buffer and are scattered into the ghost columns over the FULL y-z extent, ghost planes included (with the
// stale values the sender had there, exactly as the reference's x-exchange does, src/bound.f90:602-607); the
// z-push of a remote neighbour then overwrites the edge cells (x-ghost, z-ghost) with the right values.  With one
// "ready" for all three directions a fast z-neighbour could push BEFORE the x-scatter, which then overwrote the edge
// with stale data: seen with 32^3 boxes on a 2 x 1 x 2 process grid (errors of 1e-3 at the z-interfaces, gone under
// CUDA_LAUNCH_BLOCKING); the reference's blocking MPI_SENDRECV per direction implies the same order.
void Exchange::run_all(Ctx &ctx) const {
  for (int d = 0; d < 3; ++d) run(ctx, d);
}

void FaceFill::run(Ctx &ctx) const {
  for (int d = 0; d < 3; ++d) launch_face_ops(ctx, ops[d]);
}

}  // namespace snac
