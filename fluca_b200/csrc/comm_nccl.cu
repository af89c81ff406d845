// comm_nccl.cu -- NCCL communicator of the z-slab partition (one process per GPU, NVLink/NVSwitch).
//
// The data path has no bulk collective: slabs only exchange one ghost plane per face per field
// (grouped ncclSend/ncclRecv straight out of / into the field arrays -- planes are contiguous
// because z is the slowest index) and sum a handful of doubles per Krylov iteration
// (ncclAllReduce on the reduction buffer).  Everything is enqueued on the solver's stream.
#ifndef FLUCA_HOSTEMU
#include "solver.h"
#include <nccl.h>

namespace fluca {

#define FL_NCCL(call) \
  do { \
    ncclResult_t r_ = (call); \
    if (r_ != ncclSuccess) throw Error(FL_ERR_NCCL, std::string("NCCL error: ") + ncclGetErrorString(r_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

struct NcclComm : public Comm {
  ncclComm_t comm = nullptr;
  bool       capturable() const override { return true; } // NCCL operations are stream-ordered enqueues
  ~NcclComm() override
  {
    if (comm) ncclCommDestroy(comm);
  }
  void halo(Exec &ex, double *const *fields, int nf, long plane, int nzl, bool periodic) override
  {
    const int down = rank > 0 ? rank - 1 : (periodic ? nranks - 1 : -1);
    const int up   = rank < nranks - 1 ? rank + 1 : (periodic ? 0 : -1);
    if (nranks == 1) {
      if (!periodic) return;
      for (int f = 0; f < nf; ++f) {
        double *a = fields[f];
        copy_d2d(ex, a, a + plane * nzl, sizeof(double) * plane);
        copy_d2d(ex, a + plane * (nzl + 1), a + plane, sizeof(double) * plane);
      }
      ex.stats.launches += 2 * nf;
      return;
    }
    FL_NCCL(ncclGroupStart());
    // Post order matters when both neighbours are the same rank (2 ranks, periodic z): NCCL pairs the
    // k-th send to a peer with that peer's k-th receive from us.  Sends: plane 0 (down), plane nzl-1 (up);
    // receives: ghost nzl (their plane 0), ghost -1 (their plane nzl-1).
    for (int f = 0; f < nf; ++f) {
      double *a = fields[f];
      if (down >= 0) FL_NCCL(ncclSend(a + plane, plane, ncclDouble, down, comm, ex.stream));
      if (up >= 0) FL_NCCL(ncclSend(a + plane * nzl, plane, ncclDouble, up, comm, ex.stream));
      if (up >= 0) FL_NCCL(ncclRecv(a + plane * (nzl + 1), plane, ncclDouble, up, comm, ex.stream));
      if (down >= 0) FL_NCCL(ncclRecv(a, plane, ncclDouble, down, comm, ex.stream));
    }
    FL_NCCL(ncclGroupEnd());
    ex.stats.launches++;
  }
  void sendrecv(Exec &ex, const double *send_down, double *recv_down, const double *send_up, double *recv_up, long count, bool periodic) override
  {
    if (count <= 0) return;
    const int down = rank > 0 ? rank - 1 : (periodic ? nranks - 1 : -1);
    const int up   = rank < nranks - 1 ? rank + 1 : (periodic ? 0 : -1);
    if (nranks == 1) {
      if (!periodic) return;
      copy_d2d(ex, recv_down, send_up, sizeof(double) * count);
      copy_d2d(ex, recv_up, send_down, sizeof(double) * count);
      return;
    }
    FL_NCCL(ncclGroupStart()); // same post order as halo(): it matters when both neighbours are one rank
    if (down >= 0) FL_NCCL(ncclSend(send_down, count, ncclDouble, down, comm, ex.stream));
    if (up >= 0) FL_NCCL(ncclSend(send_up, count, ncclDouble, up, comm, ex.stream));
    if (up >= 0) FL_NCCL(ncclRecv(recv_up, count, ncclDouble, up, comm, ex.stream));
    if (down >= 0) FL_NCCL(ncclRecv(recv_down, count, ncclDouble, down, comm, ex.stream));
    FL_NCCL(ncclGroupEnd());
    ex.stats.launches++;
  }
  void allsum(Exec &ex, double *dev, int n) override
  {
    if (nranks == 1) return;
    FL_NCCL(ncclAllReduce(dev, dev, n, ncclDouble, ncclSum, comm, ex.stream));
    ex.stats.launches++;
  }
  void allgather(Exec &ex, const double *send, double *recv, long count) override
  {
    FL_NCCL(ncclAllGather(send, recv, count, ncclDouble, comm, ex.stream));
    ex.stats.launches++;
  }
};

Comm *make_nccl_comm(const void *unique_id, int id_bytes, int rank, int nranks)
{
  if (id_bytes != (int)sizeof(ncclUniqueId)) throw Error(FL_ERR_ARG, "unique id has the wrong size");
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  NcclComm *c = new NcclComm;
  c->rank = rank, c->nranks = nranks;
  ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    delete c;
    throw Error(FL_ERR_NCCL, std::string("ncclCommInitRank failed: ") + ncclGetErrorString(r));
  }
  return c;
}

int nccl_unique_id(void *out, int bytes)
{
  if (bytes < (int)sizeof(ncclUniqueId)) throw Error(FL_ERR_ARG, "unique id buffer too small");
  ncclUniqueId id;
  FL_NCCL(ncclGetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return (int)sizeof(id);
}

} // namespace fluca
#endif
