// Peer communication between the processes (one per GPU) of one NVLink/NVSwitch node.
//
// Replaces the MPI layer of the reference on the hot path: MPI_SENDRECV halos (src/bound.f90:587-650),
// the MPI_ALLTOALLW slab transpose (src/solver.f90:1244-1371) and the scalar MPI_ALLREDUCEs
// (src/chkdt.f90:48,53, src/chkdiv.f90:42-43).  Every process allocates one arena, publishes a CUDA
// IPC handle and the table of arrays in it, and maps the arenas of its peers: halos are then pushed
// straight into the neighbour's ghost cells, and the transposes of the solver are written to / read
// from the owner's slab, by ordinary kernels over NVLink.  Ordering uses per-channel sequence flags
// in the arenas, written with system-scope stores and polled by a stream-ordered kernel, so the
// host never takes part.  The bootstrap (exchange of the blobs) is done by the caller with whatever
// it has: torch.distributed here, MPI_Allgather in a Fortran host.
#include <cstring>

#include "common.h"

namespace snac {

void PeerComm::init(int nproc_, int me_, size_t bytes) {
  nproc = nproc_; me = me_;
  if (nproc > MAXPROC) throw Error("too many processes for one node");
  cap = HEAD_BYTES + bytes + 256;
  SNAC_CUDA(cudaMalloc(&arena, cap));
  SNAC_CUDA(cudaMemset(arena, 0, cap));
  used = HEAD_BYTES;
  peer[me] = arena;
  recs[me].clear();
}

double *PeerComm::alloc(size_t elems, int id, int key) {
  used = (used + 255) / 256 * 256;
  const size_t bytes = elems * sizeof(double);
  if (used + bytes > cap) throw Error("communication arena exhausted");
  double *p = reinterpret_cast<double *>(arena + used);
  recs[me].push_back(ArenaRec{id, key, (long long)used});
  used += bytes;
  return p;
}

void PeerComm::export_blob(CommBlob &b) const {
  std::memset(&b, 0, sizeof(b));
  b.proc = me;
  b.nrec = (int)recs[me].size();
  if (b.nrec > MAXREC) throw Error("too many exported arrays");
  SNAC_CUDA(cudaIpcGetMemHandle(&b.handle, arena));
  for (int i = 0; i < b.nrec; ++i) b.rec[i] = recs[me][i];
}

void PeerComm::connect(const CommBlob *blobs, int n) {
  if (n != nproc) throw Error("snac_comm_connect: expected one blob per process");
  for (int p = 0; p < nproc; ++p) {
    const CommBlob &b = blobs[p];
    if (b.proc != p) throw Error("snac_comm_connect: blobs must be ordered by process");
    if (p == me) continue;
    void *ptr = nullptr;
    SNAC_CUDA(cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess));
    peer[p] = static_cast<char *>(ptr);
    recs[p].assign(b.rec, b.rec + b.nrec);
  }
  connected = true;
}

double *PeerComm::remote(int proc, int id, int key) const {
  for (const ArenaRec &r : recs[proc])
    if (r.id == id && r.key == key) return reinterpret_cast<double *>(peer[proc] + r.off);
  throw Error("array (" + std::to_string(id) + "," + std::to_string(key) + ") is not published by process " + std::to_string(proc));
}

// error path: tell every peer to stop waiting for this process
void PeerComm::signal_abort() {
  if (!connected) return;
  const long long one = 1;
  for (int p = 0; p < nproc; ++p)
    if (p != me && peer[p]) cudaMemcpy(peer[p] + ABORT_OFFSET, &one, sizeof(one), cudaMemcpyHostToDevice);
}

// 0: fine; 1: a peer aborted; 2: a peer never arrived (timeout).  Call after a stream synchronisation.
int PeerComm::error_state() const {
  if (!connected) return 0;
  long long w[2] = {0, 0};
  if (cudaMemcpy(w, arena + ABORT_OFFSET, sizeof(w), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return (int)w[1];
}

void PeerComm::release() {
  for (int p = 0; p < nproc; ++p)
    if (p != me && peer[p]) cudaIpcCloseMemHandle(peer[p]);
  if (arena) cudaFree(arena);
  arena = nullptr;
}

struct HandshakeArgs {
  long long *remote_flag[MAXPROC];   // my slot in each peer's flag table (channel row already applied)
  volatile long long *local_flag;    // my own table row: local_flag[src]
  volatile long long *abort_word;    // my arena: [0] set by a failing peer, [1] error raised here (read by the host)
  unsigned mask;
  long long seq;
};

__device__ inline unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// A peer that throws (any error path of the C ABI) sets the abort word of every arena; a peer that dies
// never arrives: the wait is bounded (HANDSHAKE_TIMEOUT_S, far beyond any legitimate skew between the
// processes).  Either way the kernel leaves the loop and raises the error word, which the host reads at
// its next synchronisation point (snac_sync, chkdt/chkdiv) -- instead of hanging inside a kernel forever.
constexpr unsigned long long HANDSHAKE_TIMEOUT_S = 900;
__global__ void handshake_kernel(const HandshakeArgs a) {
  const int t = threadIdx.x;
  if (t >= MAXPROC || !((a.mask >> t) & 1u)) return;
  __threadfence_system();                       // everything this GPU wrote before is visible first
  *reinterpret_cast<volatile long long *>(a.remote_flag[t]) = a.seq;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (a.local_flag[t] < a.seq) {
    __nanosleep(64);
    if ((++spins & 1023u) == 0) {
      if (a.abort_word[0] != 0) { a.abort_word[1] = 1; break; }
      if (global_ns() - t0 > HANDSHAKE_TIMEOUT_S * 1000000000ull) { a.abort_word[1] = 2; break; }
    }
  }
  __threadfence_system();
}

void PeerComm::handshake(Ctx &ctx, int ch, unsigned mask) {
  ++seq[ch];
  mask &= ~(1u << me);
  mask &= (nproc >= 32) ? ~0u : ((1u << nproc) - 1u);
  if (nproc == 1 || mask == 0) return;
  if (!connected) throw Error("peer communication used before snac_comm_connect");
  HandshakeArgs a;
  for (int p = 0; p < MAXPROC; ++p) {
    a.remote_flag[p] = nullptr;
    if (p < nproc && ((mask >> p) & 1u))
      a.remote_flag[p] = reinterpret_cast<long long *>(peer[p]) + (size_t)ch * MAXPROC + me;
  }
  a.local_flag = reinterpret_cast<volatile long long *>(arena) + (size_t)ch * MAXPROC;
  a.abort_word = reinterpret_cast<volatile long long *>(arena + ABORT_OFFSET);
  a.mask = mask;
  a.seq = seq[ch];
  handshake_kernel<<<1, 32, 0, ctx.stream>>>(a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

}  // namespace snac
