// group_exchange.h — comm-buffer layout and launchers of the device-side GROUP BY exchange (group_exchange.cu).
#pragma once
#include <memory>
#include <mutex>
#include <vector>

#include "batch.h"

namespace ark {

constexpr int GX_MAX_WORLD = 16;
constexpr size_t GX_HEADER_BYTES = 4096;
constexpr unsigned long long GX_POISON = 1ull << 63;    // the source holds a key that cannot travel inline (longer than 12 bytes)
constexpr unsigned long long GX_OVERFLOW = 1ull << 62;  // the source had more records for this rank than a region holds
constexpr unsigned long long GX_COUNT_MASK = (1ull << 62) - 1;

struct GxHeader {
  unsigned long long ready[GX_MAX_WORLD];
  unsigned long long ack[GX_MAX_WORLD];
  unsigned long long count[2][GX_MAX_WORLD];
  unsigned int cursor[GX_MAX_WORLD];
  unsigned int done, flags;
};
static_assert(sizeof(GxHeader) <= GX_HEADER_BYTES, "header fits");

struct GxPeers { uint8_t* p[GX_MAX_WORLD]; };

// One rank's end of the exchange: its comm buffer and the peers' buffers as mapped into this process.
struct DistCtx {
  int rank = 0, world = 1, device = 0;
  uint8_t* comm = nullptr;
  size_t comm_bytes = 0, region_bytes = 0;
  GxPeers peers{};
  std::vector<void*> opened;      // cudaIpcOpenMemHandle results (closed on destroy)
  bool connected = false;
  unsigned long long step = 0;    // exchanges issued so far (every rank counts the same calls)
  // state carried from the push phase to the merge phase of one step
  bool pushed = false;
  // last completed step, for callers that report the exchange (bench.py): records received / groups owned here
  unsigned long long last_recv_records = 0, last_groups = 0;
  std::mutex mu;
};

void launch_exchange_push(const uint8_t* table, unsigned long long capacity, int n_acc, int key_kind, const GxPeers& peers, int world, int rank,
                          unsigned long long step, unsigned long long region_bytes, int need_count, cudaStream_t stream);
void launch_exchange_merge(uint8_t* table, unsigned long long capacity, int n_acc, const int32_t* acc_kind, const uint8_t* comm,
                           int world, int rank, unsigned long long step, unsigned long long region_bytes, unsigned int* group_count,
                           int32_t* overflow, int32_t* status, unsigned long long* total, cudaStream_t stream);
void launch_exchange_ack(const GxPeers& peers, int world, int rank, unsigned long long step, cudaStream_t stream);

}  // namespace ark
