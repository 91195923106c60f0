// Specialised lane-parallel kernel: the lane-ordered stream of one program written out as
// straight-line PTX for sm_100a. Every physical register of the stream is a real register, every
// LUT3 is one `lop3.b32` with its truth table as the immediate, inputs are `ld.global.nc` vector
// loads at first use, outputs are stored the moment they exist. K words per thread (K = 1, 2, 4):
// a warp reads/writes 32*K*4 contiguous bytes of a row per access (128-bit accesses at K = 4).
//
// The text is assembled by the CUDA driver (cuModuleLoadDataEx) for the device it runs on; tests
// assemble it with ptxas on the build host.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iterator>
#include <string>
#include <vector>

#include "device.hpp"

namespace clg {

// Live physical registers at instruction boundary `at` of the lane stream: slots whose current value was defined
// before `at` and is still read at or after it.
std::vector<uint32_t> spec_live_slots(const Flat &f, uint32_t at) {
  const FlatStream &st = f.header().stream[STREAM_LANE];
  const FlatInstr *in = f.instrs(STREAM_LANE);
  std::vector<uint8_t> defined(st.n_slots, 0), live(st.n_slots, 0);
  for (uint32_t i = 0; i < at && i < st.n_instr; i++)
    if (in[i].kind == K_LUT || in[i].kind == K_LDIN) defined[in[i].dst] = 1;
  std::vector<uint8_t> redefined(st.n_slots, 0);  // overwritten after `at` before being read: not live
  for (uint32_t i = at; i < st.n_instr; i++) {
    const FlatInstr &x = in[i];
    auto rd = [&](uint16_t sl) {
      if (defined[sl] && !redefined[sl]) live[sl] = 1;
    };
    if (x.kind == K_LUT) {
      rd(x.a);
      if (lut_nfan(x) >= 2) rd(x.b);
      if (lut_nfan(x) >= 3) rd(x.c);
    } else if (x.kind == K_STOUT)
      rd(x.a);
    if (x.kind == K_LUT || x.kind == K_LDIN) redefined[x.dst] = 1;
  }
  std::vector<uint32_t> out;
  for (uint32_t sl = 0; sl < st.n_slots; sl++)
    if (live[sl]) out.push_back(sl);
  return out;
}

// Cut points of the chunked form: n_chunks + 1 instruction indices, first 0, last n_instr. Every cut costs a store
// and a reload of each register live across it (and the reloads crowd the next kernel's register file), so each
// chunk ends, somewhere in the second half of the longest length allowed, where the fewest registers are live --
// between two instances of a flattened program that is zero. The assembler's time grows with the number of PTX
// instructions (K per LUT), so the allowed length shrinks with K.
uint32_t spec_chunk_limit(int K) {
  const uint32_t n = spec_chunk_limit();
  return std::max(2u, K == 1 ? n : K == 2 ? n / 2 : (uint32_t)((uint64_t)n * 3 / 8));
}

std::vector<uint32_t> spec_chunk_cuts(const Flat &f, int K, uint32_t limit_or_0, bool fine) {
  const FlatStream &st = f.header().stream[STREAM_LANE];
  const FlatInstr *in = f.instrs(STREAM_LANE);
  const uint32_t n = st.n_instr, limit = limit_or_0 ? std::max(2u, limit_or_0) : spec_chunk_limit(K);
  std::vector<uint32_t> cuts{0};
  if (n > limit) {
    std::vector<uint32_t> live_at(n + 1, 0);  // registers live at the boundary before instruction i
    std::vector<uint8_t> live(st.n_slots, 0);
    uint32_t count = 0;
    for (uint32_t i = n; i-- > 0;) {
      const FlatInstr &x = in[i];
      if ((x.kind == K_LUT || x.kind == K_LDIN) && live[x.dst]) live[x.dst] = 0, count--;
      auto rd = [&](uint16_t sl) {
        if (!live[sl]) live[sl] = 1, count++;
      };
      if (x.kind == K_LUT) {
        rd(x.a);
        if (lut_nfan(x) >= 2) rd(x.b);
        if (lut_nfan(x) >= 3) rd(x.c);
      } else if (x.kind == K_STOUT)
        rd(x.a);
      live_at[i] = count;
    }
    // fine: the first boundary no register crosses, if one is in reach -- the instances of a flattened program then
    // get a kernel each (run side by side, blockIdx.y) instead of one kernel for as many as fit the limit: the
    // shorter the straight-line code, the better it sits in the instruction cache (12 x mul32 at 2^24 vectors: 81 %
    // of the LOP3 pipe with one instance per kernel, 71 % with two, 47 % with five; tools/sweep_tp_chunk.py)
    // An instance longer than the limit is cut into equal parts at the cheapest boundary near each split point, so
    // that the parts of every instance come out alike and share their kernels too.
    const uint32_t min_len = std::min(limit / 2, 1024u);
    for (;;) {
      const uint32_t last = cuts.back();
      uint32_t lo, hi;
      if (fine) {
        uint32_t z = last + min_len + 1;
        while (z < n && live_at[z] != 0) z++;
        if (z >= n || n - z <= min_len) z = n;
        const uint32_t len = z - last;
        if (len <= limit) {
          if (z == n) break;
          cuts.push_back(z);
          continue;
        }
        const uint32_t parts = (len + limit - 1) / limit, target = last + len / parts, slack = len / parts / 8;
        lo = target - slack, hi = std::min(last + limit, target + slack);
      } else {
        if (n - last <= limit) break;
        lo = last + limit / 2 + 1, hi = last + limit;
      }
      uint32_t best = hi;
      for (uint32_t i = hi; i >= lo; i--)
        if (live_at[i] < live_at[best]) best = i;
      cuts.push_back(best);
    }
  }
  cuts.push_back(n);
  return cuts;
}

// Instructions [begin, end) of the lane stream in chunk-local terms: registers renumbered by first appearance and
// rows made relative to the lowest row the chunk touches in each buffer, so that two instances of the same sub-circuit
// produce EQUAL chunks and share one assembled kernel. A register that crosses a cut keeps its place in the scratch
// buffer (scr_row: its number in the whole program), so chunks that hand registers over can be shared as well when the
// instances' registers were allocated alike -- the two halves of every 64 x 64 multiplier of a flattened program are
// two kernels, not two per instance.
static SpecChunk local_code(const Flat &f, uint32_t begin, uint32_t end, std::vector<uint32_t> live_in, std::vector<uint32_t> live_out,
                            bool may_rebase) {
  const FlatHeader &h = f.header();
  const FlatInstr *in = f.instrs(STREAM_LANE);
  SpecChunk c;
  c.live_in = std::move(live_in), c.live_out = std::move(live_out);
  c.n_slots = h.stream[STREAM_LANE].n_slots;
  c.code.assign(in + begin, in + end);
  for (FlatInstr &x : c.code) {
    if (x.kind == K_LDIN && x.row >= h.n_inputs) x.row -= h.n_inputs, x.flags |= FLAG_STATE_ROW;
    if ((x.kind == K_STOUT || x.kind == K_STCONST) && x.row >= h.n_outputs) x.row -= h.n_outputs, x.flags |= FLAG_STATE_ROW;
  }
  if (!may_rebase) return c;
  uint32_t lo[3] = {UINT32_MAX, UINT32_MAX, UINT32_MAX};  // imm, out, state
  std::vector<uint16_t> name(c.n_slots, UINT16_MAX);
  uint16_t next = 0;
  auto rn = [&](uint16_t &sl) {
    if (name[sl] == UINT16_MAX) name[sl] = next++;
    sl = name[sl];
  };
  for (FlatInstr &x : c.code) {
    if (x.kind == K_LUT) rn(x.a), rn(x.b), rn(x.c), rn(x.dst);
    if (x.kind == K_LDIN) rn(x.dst);
    if (x.kind == K_STOUT) rn(x.a);
    if (x.kind != K_LUT) {
      uint32_t &l = lo[(x.flags & FLAG_STATE_ROW) ? 2 : x.kind == K_LDIN ? 0 : 1];
      l = std::min(l, x.row);
    }
  }
  c.imm_base = lo[0] == UINT32_MAX ? 0 : lo[0], c.out_base = lo[1] == UINT32_MAX ? 0 : lo[1], c.state_base = lo[2] == UINT32_MAX ? 0 : lo[2];
  for (FlatInstr &x : c.code)
    if (x.kind != K_LUT) x.row -= (x.flags & FLAG_STATE_ROW) ? c.state_base : x.kind == K_LDIN ? c.imm_base : c.out_base;
  // the live sets in local names (a register that crosses the whole chunk untouched has none: it stays in the scratch buffer)
  c.scr_row.assign(next, UINT32_MAX);
  for (std::vector<uint32_t> *lv : {&c.live_in, &c.live_out}) {
    std::vector<uint32_t> local;
    for (uint32_t sl : *lv)
      if (name[sl] != UINT16_MAX) local.push_back(name[sl]), c.scr_row[name[sl]] = sl;
    std::sort(local.begin(), local.end());
    lv->swap(local);
  }
  c.n_slots = next;
  return c;
}

bool SpecChunk::same_kernel(const SpecChunk &o) const {
  return n_slots == o.n_slots && code.size() == o.code.size() && live_in == o.live_in && live_out == o.live_out && scr_row == o.scr_row &&
         std::memcmp(code.data(), o.code.data(), code.size() * sizeof(FlatInstr)) == 0;
}

// The chain a long program runs as: cuts where few registers are live, the live sets at the cuts, chunk-local code.
std::vector<SpecChunk> spec_chunks(const Flat &f, int K, uint32_t limit, bool fine) {
  const std::vector<uint32_t> cuts = spec_chunk_cuts(f, K, limit, fine);
  const size_t n = cuts.size() - 1;
  std::vector<std::vector<uint32_t>> live(n + 1);
  for (size_t c = 1; c < n; c++) live[c] = spec_live_slots(f, cuts[c]);
  std::vector<SpecChunk> out;
  for (size_t c = 0; c < n; c++) {
    out.push_back(local_code(f, cuts[c], cuts[c + 1], live[c], live[c + 1], true));
    out.back().crosses_in = !live[c].empty(), out.back().crosses_out = !live[c + 1].empty();
  }
  return out;
}

// The launches of the chain: which chunks get assembled (one per distinct kernel) and which chunks run side by side in
// one launch; blockIdx.y then selects the instance's row bases and its part of the scratch buffer.
SpecChain spec_chain(const Flat &f, int K, uint32_t limit, bool fine) {
  SpecChain ch;
  ch.chunks = spec_chunks(f, K, limit, fine);
  const std::vector<SpecChunk> &chunks = ch.chunks;
  ch.kernel_of.resize(chunks.size());
  for (uint32_t c = 0; c < chunks.size(); c++) {
    uint32_t u = 0;
    while (u < ch.unique.size() && !chunks[c].same_kernel(chunks[ch.unique[u]])) u++;
    if (u == ch.unique.size()) ch.unique.push_back(c);
    ch.kernel_of[c] = u;
  }
  auto state_rows = [&](const SpecChunk &c) {
    std::vector<uint32_t> r;
    for (const FlatInstr &x : c.code)
      if (x.kind != K_LUT && (x.flags & FLAG_STATE_ROW)) r.push_back(c.state_base + x.row);
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
    return r;
  };
  // Groups: from a cut nothing crosses to the next such cut. Consecutive groups of the same kernels (position by
  // position) that touch disjoint carried-state rows are independent of each other and run side by side.
  const uint32_t n = (uint32_t)chunks.size();
  std::vector<uint32_t> gstart;  // first chunk of every group, then n
  for (uint32_t c = 0; c < n; c++)
    if (c == 0 || !chunks[c].crosses_in) gstart.push_back(c);
  gstart.push_back(n);
  auto group_rows = [&](uint32_t g) {
    std::vector<uint32_t> r;
    for (uint32_t c = gstart[g]; c < gstart[g + 1]; c++) {
      const std::vector<uint32_t> rc = state_rows(chunks[c]);
      r.insert(r.end(), rc.begin(), rc.end());
    }
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
    return r;
  };
  auto same_group = [&](uint32_t a, uint32_t b) {
    const uint32_t len = gstart[a + 1] - gstart[a];
    if (gstart[b + 1] - gstart[b] != len) return false;
    for (uint32_t i = 0; i < len; i++)
      if (ch.kernel_of[gstart[a] + i] != ch.kernel_of[gstart[b] + i]) return false;
    return true;
  };
  const uint32_t n_groups = (uint32_t)gstart.size() - 1;
  for (uint32_t g = 0; g < n_groups;) {
    const uint32_t len = gstart[g + 1] - gstart[g];
    std::vector<uint32_t> rows_of_run = group_rows(g);  // absolute state rows touched by the groups of the current run
    uint32_t run = 1;
    while (g + run < n_groups && run < 65535 && same_group(g, g + run)) {
      std::vector<uint32_t> rows = group_rows(g + run), both, all;
      std::set_intersection(rows.begin(), rows.end(), rows_of_run.begin(), rows_of_run.end(), std::back_inserter(both));
      if (!both.empty()) break;
      std::merge(rows.begin(), rows.end(), rows_of_run.begin(), rows_of_run.end(), std::back_inserter(all));
      rows_of_run.swap(all);
      run++;
    }
    for (uint32_t p = 0; p < len; p++) ch.launches.push_back({gstart[g] + p, run, run > 1 ? len : 1});
    g += run;
  }
  return ch;
}

static std::string emit_spec(const Flat &f, int K, bool multi_step, const SpecChunk &cc, bool chunk);

std::string emit_spec_ptx(const Flat &f, int K, bool multi_step) {
  const uint32_t n = f.header().stream[STREAM_LANE].n_instr;
  if (n == 0) fail(CLG_E_UNSUPPORTED, "no lane-ordered stream in this flat buffer");
  return emit_spec(f, K, multi_step, local_code(f, 0, n, {}, {}, false), false);
}

// One chunk of a long program as its own kernel `clg_spec_chunk`: the registers that are live across the chunk's
// cuts travel through a scratch buffer in global memory ([slot][stride] words, coalesced like rows). The caller
// advances the imm / out / state pointers by the chunk's row bases.
std::string emit_spec_chunk_ptx(const Flat &f, int K, const SpecChunk &c) { return emit_spec(f, K, false, c, true); }

static std::string emit_spec(const Flat &f, int K, bool multi_step, const SpecChunk &cc, bool chunk) {
  const FlatHeader &h = f.header();
  const FlatStream &st = h.stream[STREAM_LANE];
  if (K != 1 && K != 2 && K != 4) fail(CLG_E_INVALID, "K must be 1, 2 or 4");
  const FlatInstr *in = cc.code.data();
  const uint32_t begin = 0, end = (uint32_t)cc.code.size(), n_slots = cc.n_slots;
  const std::vector<uint32_t> *live_in = &cc.live_in, *live_out = &cc.live_out;
  std::vector<uint8_t> need_load;
  std::vector<uint32_t> store_at;
  std::string s;
  s.reserve((size_t)(end - begin) * (size_t)(40 * K + 60) + 4096);
  char buf[512];
  auto R = [&](uint32_t slot, int k) { return "%s" + std::to_string(slot * K + k); };
  auto vec = [&](uint32_t slot) {
    std::string v = K == 1 ? "" : "{";
    for (int k = 0; k < K; k++) v += (k ? "," : "") + R(slot, k);
    return v + (K == 1 ? "" : "}");
  };
  const char *vt = K == 4 ? ".v4" : K == 2 ? ".v2" : "";
  s += "//\n// generated by complogic_b200: specialised NAND-VM evaluator, " + std::to_string(st.n_instr) +
       " instructions, " + std::to_string(n_slots) + " registers x " + std::to_string(K) + " words\n//\n";
  s += ".version 8.7\n.target sm_100a\n.address_size 64\n\n";
  // multi_step: the same straight-line body inside a loop over n_steps run() calls. Carried registers stay in
  // REGISTERS between steps (loaded from `state` once before the loop, stored once after it); every step reads its
  // own immediates block and writes its own outputs block (p_imm_step / p_out_step bytes apart).
  s += std::string(".visible .entry ") + (multi_step ? "clg_spec_steps" : chunk ? "clg_spec_chunk" : "clg_spec") +
       "(\n  .param .u64 p_imm,\n  .param .u64 p_out,\n  .param .u64 p_state,\n"
       "  .param .u32 p_stride,\n  .param .u32 p_groups" +
       (multi_step ? ",\n  .param .u32 p_steps,\n  .param .u64 p_imm_step,\n  .param .u64 p_out_step" : "") +
       (chunk ? ",\n  .param .u64 p_scratch,\n  .param .u64 p_bases" : "") + "\n)\n{\n";
  std::snprintf(buf, sizeof buf,
                "  .reg .b32 %%s<%u>;\n  .reg .b32 %%t<%d>;\n  .reg .b32 %%rtid, %%rcta, %%rntid, %%gid, %%stride, %%groups;\n"
                "  .reg .b64 %%imm, %%out, %%state, %%off, %%sb, %%addr;\n  .reg .pred %%p;\n\n",
                n_slots * K + 1, K + 1);
  s += buf;
  auto C = [&](uint32_t k_state, int k) { return "%c" + std::to_string(k_state * K + k); };
  if (multi_step) {
    std::snprintf(buf, sizeof buf, "  .reg .b32 %%c<%u>;\n  .reg .b32 %%step, %%steps;\n  .reg .b64 %%immstep, %%outstep;\n  .reg .pred %%q;\n",
                  h.n_state * K + 1);
    s += buf;
  }
  s += "  ld.param.u64 %imm, [p_imm];\n  ld.param.u64 %out, [p_out];\n  ld.param.u64 %state, [p_state];\n"
       "  ld.param.u32 %stride, [p_stride];\n  ld.param.u32 %groups, [p_groups];\n"
       "  mov.u32 %rtid, %tid.x;\n  mov.u32 %rcta, %ctaid.x;\n  mov.u32 %rntid, %ntid.x;\n"
       "  mad.lo.u32 %gid, %rcta, %rntid, %rtid;\n  setp.ge.u32 %p, %gid, %groups;\n  @%p bra DONE;\n"
       "  cvta.to.global.u64 %imm, %imm;\n  cvta.to.global.u64 %out, %out;\n  cvta.to.global.u64 %state, %state;\n";
  std::snprintf(buf, sizeof buf,
                "  mul.wide.u32 %%off, %%gid, %d;\n  add.u64 %%imm, %%imm, %%off;\n  add.u64 %%out, %%out, %%off;\n"
                "  add.u64 %%state, %%state, %%off;\n  mul.wide.u32 %%sb, %%stride, 4;\n\n",
                K * 4);
  s += buf;
  if (chunk) {
    // The scratch buffer is the library's own, so its layout is chosen for cheap addresses: blocked by warp,
    // [column group / 32][register][32 lanes][K words]. A register's row is then a compile-time offset from one base per
    // thread ([%scr+imm]: no address arithmetic per access, against three instructions for a row of the caller's
    // [row][stride] buffers), a warp's access is one contiguous 128 K bytes, and a warp's whole hand-over is one
    // contiguous block of the buffer.
    std::snprintf(buf, sizeof buf,
                  "  .reg .b64 %%scr, %%scrw;\n  .reg .b32 %%wid, %%lid;\n  ld.param.u64 %%scr, [p_scratch];\n  cvta.to.global.u64 %%scr, %%scr;\n"
                  "  shr.u32 %%wid, %%gid, 5;\n  and.b32 %%lid, %%gid, 31;\n  mul.wide.u32 %%scrw, %%wid, %u;\n"
                  "  mad.wide.u32 %%scrw, %%lid, %d, %%scrw;\n  add.u64 %%scr, %%scr, %%scrw;\n",
                  h.stream[STREAM_LANE].n_slots * 128u * (uint32_t)K, K * 4);
    s += buf;
    // Row bases: blockIdx.y picks one entry {imm, out, state, -} of the launch's table, so that equal chunks (the
    // instances of a flattened program) run side by side in ONE launch, each on its own rows.
    s += "  .reg .b64 %bases, %row;\n  .reg .b32 %inst, %b<4>;\n  ld.param.u64 %bases, [p_bases];\n  cvta.to.global.u64 %bases, %bases;\n"
         "  mov.u32 %inst, %ctaid.y;\n  mad.wide.u32 %bases, %inst, 16, %bases;\n  ld.global.nc.v4.u32 {%b0,%b1,%b2,%b3}, [%bases];\n"
         "  cvt.u64.u32 %row, %b0;\n  mad.lo.u64 %imm, %sb, %row, %imm;\n  cvt.u64.u32 %row, %b1;\n  mad.lo.u64 %out, %sb, %row, %out;\n"
         "  cvt.u64.u32 %row, %b2;\n  mad.lo.u64 %state, %sb, %row, %state;\n";
    // the launch's fourth table word: which part of the scratch buffer this instance hands its registers over in
    // (groups of chunks that run side by side, position by position; a part holds every register of every column group)
    std::snprintf(buf, sizeof buf,
                  "  add.u32 %%wid, %%groups, 31;\n  shr.u32 %%wid, %%wid, 5;\n  mul.wide.u32 %%scrw, %%wid, %u;\n"
                  "  cvt.u64.u32 %%row, %%b3;\n  mad.lo.u64 %%scr, %%scrw, %%row, %%scr;\n",
                  h.stream[STREAM_LANE].n_slots * 128u * (uint32_t)K);
    s += buf;
    // Hand-over: a live-in register is reloaded from the scratch buffer right before its first use here (not at
    // the top, where hundreds of reloads would all be live at once), a live-out one is stored right after its last
    // definition here. A register that is live across the whole chunk untouched is already in the scratch buffer.
    need_load.assign(n_slots, 0), store_at.assign(n_slots, UINT32_MAX);
    for (uint32_t sl : *live_in) need_load[sl] = 1;
    std::vector<uint8_t> out_live(n_slots, 0);
    for (uint32_t sl : *live_out) out_live[sl] = 1;
    for (uint32_t i = begin; i < end; i++)
      if ((in[i].kind == K_LUT || in[i].kind == K_LDIN) && out_live[in[i].dst]) store_at[in[i].dst] = i;
  }
  // A live-in register's slot belongs to it from the top of the chunk (that is what live across the cut means), so
  // loading it early costs no register: reloads are issued `ahead` instructions before the first use -- far enough to
  // cover a DRAM round trip at two warps per scheduler, not all at the top, where the whole chunk would wait for them.
  // Scratch words this chunk does not write itself are read-only for the kernel (.nc: the assembler may keep the load
  // where it is put instead of ordering it against the chunk's stores).
  auto scr = [&](uint32_t sl) { return cc.scr_row.empty() ? sl : cc.scr_row[sl]; };  // a register's row of the scratch buffer
  std::vector<std::pair<uint32_t, uint32_t>> first_use;  // (instruction, slot), ascending
  size_t next_reload = 0;
  static const uint32_t ahead = std::getenv("CLG_SPEC_PREFETCH") ? (uint32_t)std::atoi(std::getenv("CLG_SPEC_PREFETCH")) : 768u;
  if (chunk) {
    std::vector<uint8_t> seen(n_slots, 0);
    for (uint32_t i = begin; i < end; i++) {
      const FlatInstr &x = in[i];
      auto use = [&](uint32_t sl) {
        if (sl < n_slots && need_load[sl] && !seen[sl]) seen[sl] = 1, first_use.push_back({i, sl});
      };
      if (x.kind == K_LUT) use(x.a), use(x.b), use(x.c);
      else if (x.kind == K_STOUT) use(x.a);
      if ((x.kind == K_LUT || x.kind == K_LDIN) && x.dst < n_slots) seen[x.dst] = 1;  // (defined here before any use: not a reload)
    }
  }
  auto reload = [&](uint32_t sl) {
    if (!chunk || !need_load[sl]) return;
    need_load[sl] = 0;
    std::snprintf(buf, sizeof buf, "  ld.global%s%s.u32 %s, [%%scr+%u];\n", store_at[sl] == UINT32_MAX ? ".nc.L1::no_allocate" : "", vt,
                  vec(sl).c_str(), scr(sl) * 128u * (uint32_t)K);
    s += buf;
  };
  auto reload_ahead = [&](uint32_t i) {
    while (next_reload < first_use.size() && first_use[next_reload].first <= i + ahead) reload(first_use[next_reload++].second);
  };
  auto hand_over = [&](uint32_t i, uint32_t sl) {
    if (!chunk) return;
    need_load[sl] = 0;
    if (store_at[sl] != i) return;
    std::snprintf(buf, sizeof buf, "  st.global%s.u32 [%%scr+%u], %s;\n", vt, scr(sl) * 128u * (uint32_t)K, vec(sl).c_str());
    s += buf;
  };
  const char *cvt = K == 4 ? ".v4" : K == 2 ? ".v2" : "";
  auto cvec = [&](uint32_t ks) {
    std::string v = K == 1 ? "" : "{";
    for (int k = 0; k < K; k++) v += (k ? "," : "") + C(ks, k);
    return v + (K == 1 ? "" : "}");
  };
  s += "  // body\n";
  if (multi_step) {
    s += "  ld.param.u32 %steps, [p_steps];\n  ld.param.u64 %immstep, [p_imm_step];\n  ld.param.u64 %outstep, [p_out_step];\n";
    for (uint32_t ks = 0; ks < h.n_state; ks++) {
      std::snprintf(buf, sizeof buf, "  mad.lo.u64 %%addr, %%sb, %u, %%state;\n  ld.global%s.u32 %s, [%%addr];\n", ks, cvt, cvec(ks).c_str());
      s += buf;
    }
    s += "  mov.u32 %step, 0;\nSTEP:\n";
  }
  for (uint32_t i = begin; i < end; i++) {
    const FlatInstr &x = in[i];
    if (chunk) reload_ahead(i);
    switch (x.kind) {
      case K_LUT:
        reload(x.a), reload(x.b), reload(x.c);  // operands a LUT does not read alias one it does
        for (int k = 0; k < K; k++) {
          std::snprintf(buf, sizeof buf, "  lop3.b32 %s, %s, %s, %s, 0x%02x;\n", R(x.dst, k).c_str(), R(x.a, k).c_str(),
                        R(x.b, k).c_str(), R(x.c, k).c_str(), x.tt);
          s += buf;
        }
        hand_over(i, x.dst);
        break;
      case K_LDIN: {
        const bool st_row = x.flags & FLAG_STATE_ROW;
        if (multi_step && st_row) {
          for (int k = 0; k < K; k++) {
            std::snprintf(buf, sizeof buf, "  mov.b32 %s, %s;\n", R(x.dst, k).c_str(), C(x.row, k).c_str());
            s += buf;
          }
          break;
        }
        // immediates are read-only for the whole launch (.nc); carried state rows are rewritten in place later
        // by this same thread, so their loads must stay ordered before those stores: plain coherent loads
        std::snprintf(buf, sizeof buf, "  mad.lo.u64 %%addr, %%sb, %u, %s;\n  ld.global%s%s.u32 %s, [%%addr];\n",
                      x.row, st_row ? "%state" : "%imm", st_row ? "" : ".nc.L1::no_allocate", vt,
                      vec(x.dst).c_str());
        s += buf;
        hand_over(i, x.dst);
        break;
      }
      case K_STOUT:
      case K_STCONST: {
        const bool st_row = x.flags & FLAG_STATE_ROW;
        if (x.kind == K_STOUT) reload(x.a);
        if (multi_step && st_row) {
          for (int k = 0; k < K; k++) {
            if (x.kind == K_STCONST)
              std::snprintf(buf, sizeof buf, "  mov.b32 %s, 0x%08x;\n", C(x.row, k).c_str(), (x.flags & 1) ? 0xFFFFFFFFu : 0u);
            else if (x.flags & 1)
              std::snprintf(buf, sizeof buf, "  not.b32 %s, %s;\n", C(x.row, k).c_str(), R(x.a, k).c_str());
            else
              std::snprintf(buf, sizeof buf, "  mov.b32 %s, %s;\n", C(x.row, k).c_str(), R(x.a, k).c_str());
            s += buf;
          }
          break;
        }
        std::string v = K == 1 ? "" : "{";
        for (int k = 0; k < K; k++) {
          if (x.kind == K_STCONST)
            std::snprintf(buf, sizeof buf, "  mov.b32 %%t%d, 0x%08x;\n", k, (x.flags & 1) ? 0xFFFFFFFFu : 0u);
          else if (x.flags & 1)
            std::snprintf(buf, sizeof buf, "  not.b32 %%t%d, %s;\n", k, R(x.a, k).c_str());
          else
            std::snprintf(buf, sizeof buf, "  mov.b32 %%t%d, %s;\n", k, R(x.a, k).c_str());
          s += buf;
          v += (k ? ",%t" : "%t") + std::to_string(k);
        }
        v += (K == 1 ? "" : "}");
        std::snprintf(buf, sizeof buf, "  mad.lo.u64 %%addr, %%sb, %u, %s;\n  st.global.L1::no_allocate%s.u32 [%%addr], %s;\n",
                      x.row, st_row ? "%state" : "%out", vt, v.c_str());
        s += buf;
        break;
      }
      default: fail(CLG_E_FORMAT, "unknown instruction kind");
    }
  }
  if (multi_step) {
    s += "  add.u64 %imm, %imm, %immstep;\n  add.u64 %out, %out, %outstep;\n  add.u32 %step, %step, 1;\n"
         "  setp.lt.u32 %q, %step, %steps;\n  @%q bra STEP;\n";
    for (uint32_t ks = 0; ks < h.n_state; ks++) {
      std::snprintf(buf, sizeof buf, "  mad.lo.u64 %%addr, %%sb, %u, %%state;\n  st.global%s.u32 [%%addr], %s;\n", ks, cvt, cvec(ks).c_str());
      s += buf;
    }
  }
  s += "DONE:\n  ret;\n}\n";
  return s;
}

}  // namespace clg
