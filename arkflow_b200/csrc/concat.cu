// concat.cu — concat_batches on the device: one segmented-copy launch for every data buffer of every
// column, one offset-rebase launch per var-len column, one bit-gather launch per nullable column.
//
// Stands in for arrow::compute::concat_batches at the reference's call sites
// (buffer/memory.rs:130, buffer/window.rs:131,159, processor/sql.rs:146, component/json.rs:54).
// Algorithmic traffic: every byte of every input buffer read once and written once
// (2 × 32 B/row for decoded S, 134 B/msg for raw Binary payloads).
#include <algorithm>

#include "engine.h"

namespace ark {

namespace {

struct CopySpan {
  const uint8_t* src;
  uint8_t* dst;
  unsigned long long bytes;
  unsigned long long chunk0;  // index of this span's first chunk in the global chunk space
};

constexpr unsigned long long CONCAT_CHUNK = 64 * 1024;  // bytes per CTA work item
constexpr int CONCAT_THREADS = 256;

// copy [src, src+n) → [dst, dst+n) with the widest access the mutual alignment allows
__device__ void copy_span(uint8_t* dst, const uint8_t* src, unsigned long long n, int tid, int nthreads) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src), b = reinterpret_cast<uintptr_t>(dst);
  if (((a ^ b) & 15) == 0) {
    unsigned long long head = (16 - (a & 15)) & 15;
    if (head > n) head = n;
    for (unsigned long long i = tid; i < head; i += nthreads) dst[i] = src[i];
    const unsigned long long body = (n - head) / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (unsigned long long i = tid; i < body; i += nthreads) d4[i] = s4[i];
    for (unsigned long long i = head + body * 16 + tid; i < n; i += nthreads) dst[i] = src[i];
  } else if (((a ^ b) & 7) == 0) {
    unsigned long long head = (8 - (a & 7)) & 7;
    if (head > n) head = n;
    for (unsigned long long i = tid; i < head; i += nthreads) dst[i] = src[i];
    const unsigned long long body = (n - head) / 8;
    const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(src + head);
    unsigned long long* d8 = reinterpret_cast<unsigned long long*>(dst + head);
    for (unsigned long long i = tid; i < body; i += nthreads) d8[i] = s8[i];
    for (unsigned long long i = head + body * 8 + tid; i < n; i += nthreads) dst[i] = src[i];
  } else if (((a ^ b) & 3) == 0) {
    unsigned long long head = (4 - (a & 3)) & 3;
    if (head > n) head = n;
    for (unsigned long long i = tid; i < head; i += nthreads) dst[i] = src[i];
    const unsigned long long body = (n - head) / 4;
    const unsigned* s4 = reinterpret_cast<const unsigned*>(src + head);
    unsigned* d4 = reinterpret_cast<unsigned*>(dst + head);
    for (unsigned long long i = tid; i < body; i += nthreads) d4[i] = s4[i];
    for (unsigned long long i = head + body * 4 + tid; i < n; i += nthreads) dst[i] = src[i];
  } else {
    for (unsigned long long i = tid; i < n; i += nthreads) dst[i] = src[i];
  }
}

__global__ void __launch_bounds__(CONCAT_THREADS) concat_copy_kernel(const CopySpan* spans, int n_spans, unsigned long long n_chunks) {
  for (unsigned long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n_spans - 1;  // last span with chunk0 <= chunk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (spans[mid].chunk0 <= chunk) lo = mid; else hi = mid - 1;
    }
    const CopySpan s = spans[lo];
    const unsigned long long off = (chunk - s.chunk0) * CONCAT_CHUNK;
    if (off >= s.bytes) continue;
    const unsigned long long n = s.bytes - off < CONCAT_CHUNK ? s.bytes - off : CONCAT_CHUNK;
    copy_span(s.dst + off, s.src + off, n, threadIdx.x, CONCAT_THREADS);
  }
}

struct OffsetSeg {
  const int32_t* offsets;  // element 0 of the segment
  long long row0;          // first output row of the segment
  long long byte_base;     // output byte position of the segment's first value
  int first;               // offsets[0] of the segment
};

__global__ void concat_offsets_kernel(const OffsetSeg* segs, int n_segs, long long total_rows, long long total_bytes, int32_t* out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > total_rows) return;
  if (r == total_rows) { out[r] = (int32_t)total_bytes; return; }
  int lo = 0, hi = n_segs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].row0 <= r) lo = mid; else hi = mid - 1;
  }
  const OffsetSeg s = segs[lo];
  out[r] = (int32_t)((long long)s.offsets[r - s.row0] - s.first + s.byte_base);
}

struct BitSeg {
  const uint8_t* bits;  // nullptr ⇒ all ones
  long long row0;
  int bit0;
};

__global__ void concat_bits_kernel(const BitSeg* segs, int n_segs, long long total_rows, uint8_t* out) {
  const long long ob = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ob >= (total_rows + 7) / 8) return;
  unsigned v = 0;
  long long r = ob * 8;
  int lo = 0, hi = n_segs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].row0 <= r) lo = mid; else hi = mid - 1;
  }
  int seg = lo;
  for (int i = 0; i < 8 && r < total_rows; ++i, ++r) {
    while (seg + 1 < n_segs && segs[seg + 1].row0 <= r) ++seg;
    const BitSeg s = segs[seg];
    const long long p = r - s.row0 + s.bit0;
    const unsigned bit = s.bits ? (s.bits[p >> 3] >> (p & 7)) & 1u : 1u;
    v |= bit << i;
  }
  out[ob] = (uint8_t)v;
}

}  // namespace

// Concatenates batches that already live in HBM.  Var-len extents must be resolved.
Batch concat_device(std::vector<Batch>& ins, cudaStream_t stream) {
  if (ins.empty()) fail(ARK_ERR_PROCESS, "Merge batches failed: no batches");
  const size_t ncol = ins[0].cols.size();
  for (size_t b = 1; b < ins.size(); ++b) {
    if (ins[b].cols.size() != ncol) fail(ARK_ERR_PROCESS, "Merge batches failed: Invalid argument error: batches have different column counts");
    for (size_t c = 0; c < ncol; ++c) {
      if (ins[b].cols[c].field.type != ins[0].cols[c].field.type || ins[b].cols[c].field.format != ins[0].cols[c].field.format)
        fail(ARK_ERR_PROCESS, "Merge batches failed: Invalid argument error: column types must match schema types, expected " +
                                  std::string(dtype_name(ins[0].cols[c].field.type)) + " but found " +
                                  dtype_name(ins[b].cols[c].field.type) + " at column index " + std::to_string(c));
    }
  }
  {
    std::vector<Column*> all;  // one round trip for the extents of every batch
    for (auto& b : ins) for (auto& c : b.cols) all.push_back(&c);
    resolve_varlen_extents_many(all, stream);
  }
  if (ins.size() == 1) return ins[0];
  int64_t total_rows = 0;
  std::vector<int64_t> row0(ins.size());
  for (size_t b = 0; b < ins.size(); ++b) { row0[b] = total_rows; total_rows += ins[b].num_rows; }

  Batch out;
  out.num_rows = total_rows;
  out.input_name = ins[0].input_name;
  std::vector<CopySpan> spans;
  std::vector<BufferPtr> keep;
  struct PendingOffsets { std::vector<OffsetSeg> segs; int32_t* out; int64_t total_bytes; };
  struct PendingBits { std::vector<BitSeg> segs; uint8_t* out; };
  std::vector<PendingOffsets> poffs;
  std::vector<PendingBits> pbits;
  auto add_span = [&](const uint8_t* src, uint8_t* dst, int64_t bytes) {
    if (bytes > 0) spans.push_back({src, dst, (unsigned long long)bytes, 0});
  };
  for (size_t c = 0; c < ncol; ++c) {
    const Column& c0 = ins[0].cols[c];
    if (!c0.present) fail(ARK_ERR_UNSUPPORTED, "concat of a column with Arrow type '" + c0.field.format + "'");
    Column oc;
    oc.field = c0.field; oc.length = total_rows;
    bool any_null = false;
    for (auto& b : ins) any_null = any_null || (b.cols[c].validity != nullptr);
    switch (c0.field.type) {
      case DType::Int64: case DType::Float64: {
        BufferPtr d = device_alloc((size_t)total_rows * 8 + 16);
        for (size_t b = 0; b < ins.size(); ++b) add_span(ins[b].cols[c].data, (uint8_t*)d.get() + row0[b] * 8, ins[b].num_rows * 8);
        oc.data = (const uint8_t*)d.get(); oc.data_bytes = total_rows * 8; oc.owners.push_back(d);
        break;
      }
      case DType::Bool: {
        BufferPtr d = device_alloc((size_t)(total_rows + 7) / 8 + 16);
        PendingBits pb; pb.out = (uint8_t*)d.get();
        for (size_t b = 0; b < ins.size(); ++b) if (ins[b].num_rows) pb.segs.push_back({ins[b].cols[c].data, row0[b], ins[b].cols[c].data_bit0});
        pbits.push_back(std::move(pb));
        oc.data = (const uint8_t*)d.get(); oc.data_bit0 = 0; oc.data_bytes = (total_rows + 7) / 8; oc.owners.push_back(d);
        break;
      }
      case DType::Utf8: case DType::Binary: {
        int64_t total_bytes = 0;
        for (auto& b : ins) total_bytes += std::max<int64_t>(b.cols[c].data_bytes, 0);
        if (total_bytes > 2147483647ll)
          fail(ARK_ERR_PROCESS, "Merge batches failed: Invalid argument error: offset overflow, concatenated Utf8 column exceeds 2 GiB");
        BufferPtr o = device_alloc((size_t)(total_rows + 1) * 4 + 16), d = device_alloc((size_t)total_bytes + 16);
        PendingOffsets po; po.out = (int32_t*)o.get(); po.total_bytes = total_bytes;
        int64_t byte_base = 0;
        for (size_t b = 0; b < ins.size(); ++b) {
          const Column& sc = ins[b].cols[c];
          if (ins[b].num_rows == 0) continue;
          po.segs.push_back({sc.offsets, row0[b], byte_base, (int)sc.first_offset});
          add_span(sc.data + sc.first_offset, (uint8_t*)d.get() + byte_base, sc.data_bytes);
          byte_base += sc.data_bytes;
        }
        poffs.push_back(std::move(po));
        oc.offsets = (const int32_t*)o.get(); oc.data = (const uint8_t*)d.get(); oc.data_bytes = total_bytes; oc.first_offset = 0;
        oc.owners = {o, d};
        break;
      }
      default: break;  // Null type: nothing to copy
    }
    if (any_null && total_rows > 0) {
      BufferPtr v = device_alloc((size_t)(total_rows + 7) / 8 + 16);
      PendingBits pb; pb.out = (uint8_t*)v.get();
      for (size_t b = 0; b < ins.size(); ++b) if (ins[b].num_rows) pb.segs.push_back({ins[b].cols[c].validity, row0[b], ins[b].cols[c].validity_bit0});
      pbits.push_back(std::move(pb));
      oc.validity = (const uint8_t*)v.get(); oc.validity_bit0 = 0; oc.null_count = -1; oc.owners.push_back(v);
    }
    out.cols.push_back(std::move(oc));
  }
  // one upload for all descriptors
  unsigned long long n_chunks = 0;
  for (auto& s : spans) { s.chunk0 = n_chunks; n_chunks += (s.bytes + CONCAT_CHUNK - 1) / CONCAT_CHUNK; }
  size_t desc_bytes = spans.size() * sizeof(CopySpan);
  for (auto& p : poffs) desc_bytes += p.segs.size() * sizeof(OffsetSeg);
  for (auto& p : pbits) desc_bytes += p.segs.size() * sizeof(BitSeg);
  BufferPtr hdesc = pinned_alloc(desc_bytes + 64), ddesc = device_alloc(desc_bytes + 64);
  size_t pos = 0;
  auto put = [&](const void* src, size_t n) -> const void* {
    memcpy((char*)hdesc.get() + pos, src, n);
    const void* dptr = (char*)ddesc.get() + pos;
    pos += n;
    return dptr;
  };
  const CopySpan* d_spans = (const CopySpan*)put(spans.data(), spans.size() * sizeof(CopySpan));
  std::vector<const OffsetSeg*> d_offs;
  std::vector<const BitSeg*> d_bits;
  for (auto& p : poffs) d_offs.push_back((const OffsetSeg*)put(p.segs.data(), p.segs.size() * sizeof(OffsetSeg)));
  for (auto& p : pbits) d_bits.push_back((const BitSeg*)put(p.segs.data(), p.segs.size() * sizeof(BitSeg)));
  if (pos) ARK_CUDA(cudaMemcpyAsync(ddesc.get(), hdesc.get(), pos, cudaMemcpyHostToDevice, stream));
  if (n_chunks) {
    KernelTimer t("concat_copy_kernel", stream);
    const unsigned grid = (unsigned)std::min<unsigned long long>(n_chunks, 148ull * 16);
    concat_copy_kernel<<<grid, CONCAT_THREADS, 0, stream>>>(d_spans, (int)spans.size(), n_chunks);
  }
  for (size_t i = 0; i < poffs.size(); ++i) {
    KernelTimer t("concat_offsets_kernel", stream);
    if (poffs[i].segs.empty()) { ARK_CUDA(cudaMemsetAsync(poffs[i].out, 0, 4, stream)); continue; }
    concat_offsets_kernel<<<(unsigned)ceil_div(total_rows + 1, 256), 256, 0, stream>>>(d_offs[i], (int)poffs[i].segs.size(), total_rows,
                                                                                       poffs[i].total_bytes, poffs[i].out);
  }
  for (size_t i = 0; i < pbits.size(); ++i) {
    if (pbits[i].segs.empty()) continue;
    KernelTimer t("concat_bits_kernel", stream);
    concat_bits_kernel<<<(unsigned)ceil_div((total_rows + 7) / 8, 256), 256, 0, stream>>>(d_bits[i], (int)pbits[i].segs.size(), total_rows, pbits[i].out);
  }
  ARK_CUDA(cudaGetLastError());
  ARK_CUDA(cudaStreamSynchronize(stream));  // descriptors live in pooled blocks: finish before they are recycled
  return out;
}

}  // namespace ark

using namespace ark;

extern "C" int ark_concat_batches(int n, ArrowArray* ins, ArrowSchema* in_schemas, ArrowArray* out, ArrowSchema* out_schema) {
  std::vector<BufferPtr> owners;
  for (int i = 0; i < n; ++i) owners.push_back(adopt_array(&ins[i]));
  try {
    if (n <= 0) fail(ARK_ERR_PROCESS, "Merge batches failed: no batches");
    StreamLease lease;
    std::vector<Batch> bs;
    for (int i = 0; i < n; ++i) bs.push_back(import_host((const ArrowArray*)owners[i].get(), &in_schemas[i], nullptr, lease.s));
    Batch r = concat_device(bs, lease.s);
    export_host(r, lease.s, out, out_schema);
    return ARK_OK;
  } catch (const ArkError& e) {
    set_last_error(e.what());
    return e.code;
  }
}

extern "C" int ark_concat_batches_device(int n, ArrowDeviceArray* ins, ArrowSchema* in_schemas, ArrowDeviceArray* out, ArrowSchema* out_schema) {
  std::vector<BufferPtr> owners;
  for (int i = 0; i < n; ++i) owners.push_back(adopt_array(&ins[i].array));
  try {
    if (n <= 0) fail(ARK_ERR_PROCESS, "Merge batches failed: no batches");
    StreamLease lease;
    std::vector<Batch> bs;
    for (int i = 0; i < n; ++i) {
      ArrowDeviceArray view = ins[i];
      view.array = *(const ArrowArray*)owners[i].get();
      bs.push_back(import_device(&view, &in_schemas[i], nullptr, owners[i]));
    }
    Batch r = concat_device(bs, lease.s);
    export_device(r, out, out_schema);
    return ARK_OK;
  } catch (const ArkError& e) {
    set_last_error(e.what());
    return e.code;
  }
}
