// engine.cu — SqlProcessor (host side) and the filter/project executor.
#include "engine.h"

#include <algorithm>

#include "json_mini.h"

namespace ark {

std::unique_ptr<SqlProcessor> SqlProcessor::from_config(const char* config_json) {
  // reference: sql.rs:235-239 — the message really says "Batch processor" (copy-paste in the reference)
  if (!config_json) fail(ARK_ERR_CONFIG, "Batch processor configuration is missing");
  JsonValue cfg = parse_json(config_json);
  if (cfg.kind == JsonValue::Null) fail(ARK_ERR_CONFIG, "Batch processor configuration is missing");
  if (cfg.kind != JsonValue::Object) fail(ARK_ERR_SERIALIZATION, "invalid type: expected struct SqlProcessorConfig");
  const JsonValue* q = cfg.get("query");
  if (!q) fail(ARK_ERR_SERIALIZATION, "missing field `query`");
  if (q->kind != JsonValue::String) fail(ARK_ERR_SERIALIZATION, "invalid type for `query`: expected a string");
  auto p = std::make_unique<SqlProcessor>();
  p->query_text = q->str;
  if (const JsonValue* t = cfg.get("table_name")) {
    if (t->kind == JsonValue::String) p->table_name = t->str;
    else if (t->kind != JsonValue::Null) fail(ARK_ERR_SERIALIZATION, "invalid type for `table_name`: expected a string");
  }
  bool resolved = false;  // the shim owns Resource.temporary: it checks the names (sql.rs:70-86) and says so
  if (const JsonValue* r = cfg.get("temporaries_resolved")) resolved = r->kind == JsonValue::Bool && r->b;
  if (const JsonValue* tl = cfg.get("temporary_list")) {
    if (tl->kind == JsonValue::Array && !tl->arr.empty() && !resolved) {
      // sql.rs:70-86: temporaries are looked up in Resource; without the shim's confirmation a configured
      // temporary is by construction "not found".
      const JsonValue* nm = tl->arr[0].get("name");
      fail(ARK_ERR_PROCESS, "Temporary " + (nm && nm->kind == JsonValue::String ? nm->str : std::string("?")) + " not found");
    }
  }
  p->ast = parse_sql(p->query_text);  // throws "SQL query error: …" (sql.rs:92-98)
  return p;
}

std::shared_ptr<const Plan> SqlProcessor::plan_for(const std::vector<Field>& fields) {
  std::string key = schema_fingerprint(fields);
  {
    std::lock_guard<std::mutex> l(mu_);
    auto it = plans_.find(key);
    if (it != plans_.end()) { if (it->second->kind == Plan::Aggregate) last_agg_ = it->second; return it->second; }
  }
  auto plan = std::make_shared<Plan>(bind_query(ast, table_name, fields));
  std::lock_guard<std::mutex> l(mu_);
  if (plans_.size() > 64) plans_.clear();
  plans_[key] = plan;
  if (plan->kind == Plan::Aggregate) last_agg_ = plan;
  return plan;
}

std::shared_ptr<const Plan> SqlProcessor::last_aggregate_plan() {
  std::lock_guard<std::mutex> l(mu_);
  return last_agg_;
}

std::shared_ptr<const Plan> SqlProcessor::join_plan_for(const std::vector<std::string>& names,
                                                        const std::vector<std::vector<Field>>& tables) {
  std::string key = "J";
  for (size_t i = 0; i < names.size(); ++i) { key += names[i]; key += '\x1d'; key += schema_fingerprint(tables[i]); key += '\x1c'; }
  {
    std::lock_guard<std::mutex> l(mu_);
    auto it = plans_.find(key);
    if (it != plans_.end()) return it->second;
  }
  auto plan = std::make_shared<Plan>(bind_join(ast, names, tables));
  std::lock_guard<std::mutex> l(mu_);
  if (plans_.size() > 64) plans_.clear();
  plans_[key] = plan;
  return plan;
}

Batch SqlProcessor::execute(const Plan& plan, Batch& in, cudaStream_t stream) {
  switch (plan.kind) {
    case Plan::FilterProject: {
      Batch r = run_filter_project(plan, in, stream);
      return plan.concats.empty() ? r : apply_concats(plan, r, stream);
    }
    case Plan::Aggregate: return run_aggregate(plan, in, stream);
    default: fail(ARK_ERR_PROCESS, "internal: join plan executed through the single-table path");
  }
}

// ---- filter / project ---------------------------------------------------------------------------------

namespace {

struct PendingOut {
  int out_index;            // index into plan.outputs
  BufferPtr data, offsets, valid_bytes;
  bool is_bool = false;
};

}  // namespace

const char* vm_error_text(int e) {
  switch (e) {
    case VMERR_DIV_ZERO: return "Arrow error: Divide by zero error";
    case VMERR_OVERFLOW: return "Arrow error: Arithmetic overflow: Overflow happened on: -9223372036854775808 / -1";
    case VMERR_OVERFLOW_MOD: return "Arrow error: Arithmetic overflow: Overflow happened on: -9223372036854775808 % -1";
    case VMERR_CAST: return "Arrow error: Cast error: Can't cast value to type Int64";
    default: return "unknown evaluation error";
  }
}


// ---- fast path: filter_project_tma.cu ------------------------------------------------------------------
int filter_project_tma_tile_rows();
int filter_project_tma_max_fixed_out();
int filter_project_tma_desc_stride();
size_t filter_project_tma_desc_words(int64_t n_tiles);
void filter_project_tma_note_avg_len(double avg);
bool launch_filter_project_tma(int64_t n_rows, const void* pred_in, int n_fixed_out, const void* const* fixed_in, void* const* fixed_out,
                               const int32_t* offsets_in, const uint8_t* data_in, int64_t data_bytes, int32_t* offsets_out, uint8_t* data_out,
                               int cmp, int is_f64, uint64_t constant, unsigned long long* desc, unsigned int* ticket, long long* totals,
                               cudaStream_t stream);

// SELECT <fixed-width cols…> [, <1 var-len col>] WHERE <fixed col> <cmp> <literal>, no NULLs in the used columns.
static bool try_fast_filter(const Plan& plan, Batch& in, Batch& out, cudaStream_t stream) {
  if (!plan.has_pred || !plan.simple.enabled || plan.outputs.empty()) return false;
  auto src_col = [&](int slot) -> Column& { return in.cols[plan.used_cols[slot]]; };
  for (size_t s = 0; s < plan.used_cols.size(); ++s) if (src_col((int)s).validity) return false;
  const int max_fixed = filter_project_tma_max_fixed_out();
  std::vector<int> fixed_outs;
  int varlen_out = -1;
  for (size_t i = 0; i < plan.outputs.size(); ++i) {
    const OutputCol& oc = plan.outputs[i];
    if (oc.src.kind != ValueSource::PassThrough) return false;
    const DType t = src_col(oc.src.slot).field.type;
    if (t == DType::Int64 || t == DType::Float64) { if ((int)fixed_outs.size() == max_fixed) return false; fixed_outs.push_back((int)i); }
    else if (t == DType::Utf8 || t == DType::Binary) { if (varlen_out >= 0) return false; varlen_out = (int)i; }
    else return false;
  }
  const int64_t n = in.num_rows;
  const int n_tiles = (int)ceil_div(n, filter_project_tma_tile_rows());
  const void* fin[8] = {nullptr};
  void* fout[8] = {nullptr};
  std::vector<BufferPtr> fbuf(fixed_outs.size());
  BufferPtr obuf, dbuf;
  for (size_t c = 0; c < fixed_outs.size(); ++c) {
    fin[c] = src_col(plan.outputs[fixed_outs[c]].src.slot).data;
    fbuf[c] = device_alloc((size_t)n * 8 + 16);
    fout[c] = fbuf[c].get();
  }
  const Column* vs = varlen_out >= 0 ? &src_col(plan.outputs[varlen_out].src.slot) : nullptr;
  if (vs) { obuf = device_alloc((size_t)(n + 1) * 4 + 16); dbuf = device_alloc((size_t)std::max<int64_t>(varlen_bytes_bound(*vs), 0) + 32); }
  const size_t desc_bytes = filter_project_tma_desc_words(n_tiles) * 8;
  const size_t scratch_bytes = round_up((int64_t)desc_bytes + 64 + FP_CHANNELS * 8, 256);
  BufferPtr scratch = device_alloc(scratch_bytes);
  ARK_CUDA(cudaMemsetAsync(scratch.get(), 0, scratch_bytes, stream));
  // the kernel's last tile stores the two totals straight into pinned host memory (device-accessible under unified
  // addressing): one stream synchronize, no separate device→host copy on the critical path of every call
  BufferPtr host_res = pinned_alloc(64);
  long long* totals = (long long*)host_res.get();
  if (!launch_filter_project_tma(n, src_col(plan.simple.slot).data, (int)fixed_outs.size(), fin, fout, vs ? vs->offsets : nullptr,
                                 vs ? vs->data : nullptr, vs ? vs->data_bytes : 0 /* -1 = unknown */, vs ? (int32_t*)obuf.get() : nullptr,
                                 vs ? (uint8_t*)dbuf.get() : nullptr, plan.simple.cmp, plan.simple.is_f64, plan.simple.constant,
                                 (unsigned long long*)scratch.get(), (unsigned int*)((char*)scratch.get() + desc_bytes), totals, stream))
    return false;
  ARK_CUDA(cudaGetLastError());
  ARK_CUDA(cudaStreamSynchronize(stream));
  const int64_t count = ((const int64_t*)host_res.get())[0], bytes = ((const int64_t*)host_res.get())[1];
  if (vs && count > 0) filter_project_tma_note_avg_len((double)bytes / (double)count);
  out.cols.resize(plan.outputs.size());
  out.num_rows = count;
  for (size_t i = 0; i < plan.outputs.size(); ++i) {
    const OutputCol& oc = plan.outputs[i];
    Column& c = out.cols[i];
    c.field.name = oc.name; c.field.type = oc.src.type; c.field.nullable = oc.src.nullable;
    c.length = count; c.null_count = 0;
    if ((int)i == varlen_out) {
      c.offsets = (const int32_t*)obuf.get(); c.data = (const uint8_t*)dbuf.get(); c.data_bytes = bytes; c.first_offset = 0;
      c.owners = {obuf, dbuf};
    } else {
      size_t k = 0;
      while (fixed_outs[k] != (int)i) ++k;
      c.data = (const uint8_t*)fbuf[k].get(); c.data_bytes = count * 8; c.owners = {fbuf[k]};
    }
  }
  return true;
}

// CAST(Binary AS Utf8) columns: arrow-cast (safe = false) fails the query on invalid UTF-8.
static void validate_utf8_outputs(const Plan& plan, Batch& out, cudaStream_t stream) {
  bool any = false;
  for (auto& oc : plan.outputs) any = any || oc.src.validate_utf8;
  if (!any || out.num_rows == 0) return;
  BufferPtr flag = device_alloc(16), h = pinned_alloc(16);
  ARK_CUDA(cudaMemsetAsync(flag.get(), 0, 16, stream));
  for (size_t i = 0; i < plan.outputs.size(); ++i) {
    if (!plan.outputs[i].src.validate_utf8) continue;
    const Column& c = out.cols[i];
    launch_utf8_validate(c.offsets, c.data, c.validity, c.validity_bit0, c.length, (int*)flag.get(), stream);
  }
  ARK_CUDA(cudaMemcpyAsync(h.get(), flag.get(), 4, cudaMemcpyDeviceToHost, stream));
  ARK_CUDA(cudaStreamSynchronize(stream));
  if (*(int*)h.get()) fail(ARK_ERR_PROCESS, "Collection query results error: Arrow error: Cast error: Cannot cast binary to string: invalid utf-8 sequence");
}

Batch run_filter_project(const Plan& plan, Batch& in, cudaStream_t stream) {
  const int64_t n = in.num_rows;
  Batch out;
  out.input_name = in.input_name;
  auto src_col = [&](int slot) -> Column& { return in.cols[plan.used_cols[slot]]; };

  if (plan.identity) {  // SELECT * FROM flow: the batch itself (zero copy)
    out.cols = in.cols;
    out.num_rows = n;
    return out;
  }

  {  // outputs are allocated worst-case: an upper bound on each var-len source is enough
    std::vector<int> unknown;
    for (auto& oc : plan.outputs)
      if (oc.src.kind == ValueSource::PassThrough) {
        const Column& c = in.cols[plan.used_cols[oc.src.slot]];
        if ((c.field.type == DType::Utf8 || c.field.type == DType::Binary) && varlen_bytes_bound(c) < 0) unknown.push_back(plan.used_cols[oc.src.slot]);
      }
    if (!unknown.empty()) resolve_varlen_extents(in, unknown, stream);
  }

  if (try_fast_filter(plan, in, out, stream)) {
    validate_utf8_outputs(plan, out, stream);
    if (plan.limit >= 0 && out.num_rows > plan.limit) {
      out.num_rows = plan.limit;
      for (auto& c : out.cols) { c.length = plan.limit; if (c.field.type == DType::Utf8 || c.field.type == DType::Binary) c.data_bytes = -1; }
      std::vector<int> all;
      for (size_t i = 0; i < out.cols.size(); ++i) all.push_back((int)i);
      resolve_varlen_extents(out, all, stream);
    }
    return out;
  }

  out.cols.resize(plan.outputs.size());
  std::vector<int> todo;  // outputs that need the kernel
  for (size_t i = 0; i < plan.outputs.size(); ++i) {
    const OutputCol& oc = plan.outputs[i];
    Column& c = out.cols[i];
    c.field.name = oc.name; c.field.type = oc.src.type; c.field.nullable = oc.src.nullable;
    if (!plan.has_pred && oc.src.kind == ValueSource::PassThrough) {
      Column& s = src_col(oc.src.slot);  // no filter: the projected column is the input column
      Field f = c.field;
      c = s; c.field = f;
    } else todo.push_back((int)i);
  }
  out.num_rows = n;

  int64_t count = n;
  if (!todo.empty() || plan.has_pred) {
    const int n_tiles = (int)ceil_div(n, FP_TILE);
    if (n_tiles <= 0) fail(ARK_ERR_PROCESS, "internal: empty batch reached the filter kernel");
    std::vector<PendingOut> pend;
    // scratch: [desc | ticket | totals | error] per launch; results gathered into one pinned block
    const size_t desc_bytes = (size_t)n_tiles * FP_CHANNELS * 8;
    const size_t scratch_bytes = round_up((int64_t)desc_bytes + 64 + FP_CHANNELS * 8, 256);
    size_t pos = 0;
    std::vector<BufferPtr> scratches;
    std::vector<int> launch_n_out;
    BufferPtr host_res = pinned_alloc(4096);
    int launches = 0;
    bool first = true;
    while (first || pos < todo.size()) {
      first = false;
      FpParams P;
      memset(&P, 0, sizeof P);
      P.n_rows = n; P.n_tiles = n_tiles;
      for (size_t s = 0; s < plan.used_cols.size(); ++s) P.cols[s] = src_col((int)s).view();
      int pred_kind = !plan.has_pred ? 0 : (plan.simple.enabled ? 1 : 2);
      if (pred_kind == 1) { P.sp_slot = plan.simple.slot; P.sp_cmp = plan.simple.cmp; P.sp_is_f64 = plan.simple.is_f64; P.sp_const = plan.simple.constant; }
      if (pred_kind == 2) P.pred = plan.pred;
      int n_progs = 0;
      while (pos < todo.size() && P.n_out < FP_MAX_OUT) {
        const int oi = todo[pos];
        const OutputCol& oc = plan.outputs[oi];
        const bool computed = oc.src.kind == ValueSource::Computed;
        const bool varlen = !computed && (oc.src.type == DType::Utf8 || oc.src.type == DType::Binary);
        if (computed && n_progs == FP_MAX_PROGS) break;
        if (varlen && P.n_varlen == FP_MAX_VARLEN) break;
        FpOutput& fo = P.outs[P.n_out];
        PendingOut po; po.out_index = oi;
        const Column* s = computed ? nullptr : &src_col(oc.src.slot);
        bool want_valid = computed ? oc.src.nullable : (s->validity != nullptr);
        if (want_valid) { po.valid_bytes = device_alloc((size_t)n); fo.out_valid = (uint8_t*)po.valid_bytes.get(); fo.write_validity = 1; }
        if (computed) {
          fo.prog = n_progs; P.progs[n_progs++] = oc.src.prog;
          if (oc.src.type == DType::Bool) { fo.kind = FP_OUT_COMPUTED_BOOL; po.is_bool = true; po.data = device_alloc((size_t)n); }
          else { fo.kind = FP_OUT_COMPUTED8; po.data = device_alloc((size_t)n * 8); }
        } else if (varlen) {
          fo.kind = FP_OUT_VARLEN; fo.slot = oc.src.slot; fo.varlen_idx = P.n_varlen;
          P.varlen_slot[P.n_varlen++] = oc.src.slot;
          po.offsets = device_alloc((size_t)(n + 1) * 4);
          po.data = device_alloc((size_t)std::max<int64_t>(varlen_bytes_bound(*s), 0) + 16);
          fo.out_offsets = (int32_t*)po.offsets.get();
        } else if (oc.src.type == DType::Bool) {
          fo.kind = FP_OUT_BOOL; fo.slot = oc.src.slot; po.is_bool = true; po.data = device_alloc((size_t)n);
        } else if (oc.src.type == DType::Int64 || oc.src.type == DType::Float64) {
          fo.kind = FP_OUT_FIXED8; fo.slot = oc.src.slot; po.data = device_alloc((size_t)n * 8);
        } else {
          fail(ARK_ERR_UNSUPPORTED, std::string("projection of a ") + dtype_name(oc.src.type) + " column");
        }
        fo.out_data = po.data.get();
        pend.push_back(std::move(po));
        ++P.n_out; ++pos;
      }
      BufferPtr scratch = device_alloc(scratch_bytes);
      ARK_CUDA(cudaMemsetAsync(scratch.get(), 0, scratch_bytes, stream));
      P.desc = (unsigned long long*)scratch.get();
      P.ticket = (unsigned int*)((char*)scratch.get() + desc_bytes);
      P.totals = (long long*)((char*)scratch.get() + desc_bytes + 16);
      P.error = (int32_t*)((char*)scratch.get() + desc_bytes + 8);
      launch_filter_project(P, pred_kind, stream);
      ARK_CUDA(cudaGetLastError());
      if (launches >= 32) fail(ARK_ERR_UNSUPPORTED, "too many output columns");
      // totals[FP_CHANNELS] + error live contiguously from desc_bytes+8: copy 8 + 8 + 24 bytes
      ARK_CUDA(cudaMemcpyAsync((char*)host_res.get() + launches * 64, (char*)scratch.get() + desc_bytes, 64,
                               cudaMemcpyDeviceToHost, stream));
      scratches.push_back(scratch);
      launch_n_out.push_back(P.n_out);
      ++launches;
    }
    ARK_CUDA(cudaStreamSynchronize(stream));
    for (int l = 0; l < launches; ++l) {
      const char* base = (const char*)host_res.get() + l * 64;
      int32_t err = *(const int32_t*)(base + 8);
      if (err) fail(ARK_ERR_PROCESS, std::string("Collection query results error: ") + vm_error_text(err));
    }
    count = *(const int64_t*)((const char*)host_res.get() + 16);
    // finish the columns produced by the kernel
    size_t pi = 0;
    for (int l = 0; l < launches; ++l) {
      const int64_t* totals = (const int64_t*)((const char*)host_res.get() + l * 64 + 16);
      int vseen = 0;
      const int n_out = launch_n_out[l];
      for (int k = 0; k < n_out; ++k, ++pi) {
        PendingOut& po = pend[pi];
        const OutputCol& oc = plan.outputs[po.out_index];
        Column& c = out.cols[po.out_index];
        c.length = count; c.present = true;
        const bool varlen = po.offsets != nullptr;
        if (varlen) {
          c.offsets = (const int32_t*)po.offsets.get(); c.data = (const uint8_t*)po.data.get();
          c.data_bytes = totals[1 + vseen]; c.first_offset = 0; ++vseen;
          c.owners = {po.offsets, po.data};
        } else if (po.is_bool) {
          BufferPtr bits = device_alloc((size_t)(count + 7) / 8 + 1);
          launch_pack_bits((const uint8_t*)po.data.get(), count, (uint8_t*)bits.get(), nullptr, stream);
          c.data = (const uint8_t*)bits.get(); c.data_bit0 = 0; c.data_bytes = (count + 7) / 8;
          c.owners = {bits, po.data};
        } else {
          c.data = (const uint8_t*)po.data.get(); c.data_bytes = count * 8; c.owners = {po.data};
        }
        if (po.valid_bytes && count > 0) {
          BufferPtr bits = device_alloc((size_t)(count + 7) / 8 + 1);
          launch_pack_bits((const uint8_t*)po.valid_bytes.get(), count, (uint8_t*)bits.get(), nullptr, stream);
          c.validity = (const uint8_t*)bits.get(); c.validity_bit0 = 0; c.null_count = -1;
          c.owners.push_back(bits); c.owners.push_back(po.valid_bytes);
        } else { c.validity = nullptr; c.null_count = 0; }
        (void)oc;
      }
    }
    // zero-copy passthrough columns of a predicate-less projection keep length n (= count)
    out.num_rows = count;
  }

  validate_utf8_outputs(plan, out, stream);
  if (plan.limit >= 0 && out.num_rows > plan.limit) {
    // LIMIT k without ORDER BY on a single-partition scan: the first k surviving rows
    out.num_rows = plan.limit;
    for (auto& c : out.cols) {
      c.length = plan.limit;
      if (c.field.type == DType::Utf8 || c.field.type == DType::Binary) c.data_bytes = -1;
    }
    std::vector<int> all;
    for (size_t i = 0; i < out.cols.size(); ++i) all.push_back((int)i);
    resolve_varlen_extents(out, all, stream);
  }
  return out;
}

}  // namespace ark
