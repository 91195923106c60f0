// The compiler pass the north-star adds in front of the device: take the reference's in-order
// NAND program (`Simulation::ops`, executed by simulation.rs:16-30) and produce the flat,
// level-ordered, register-allocated device buffer (flat.hpp).
//
//   1. SSA rename.  Ops execute strictly in list order, registers may be written more than once
//      and read before they are written (simulation.rs:16-30). Walking the
//      list with a "current value of register r" table turns that into a pure dataflow graph.
//      The graph is an and-inverter graph with complemented edges, so `Nand(a,a,out)` (more than
//      half of what gates.rs emits) costs nothing, and structural hashing shares duplicates.
//   2. Set handling (simulation.rs:26): id < n_immediates reads input row id, otherwise the
//      literal. Reads of a never-written register see `false` (fresh Simulation, compile.rs:199)
//      or, in stateful mode, the carried value of that register.
//   3. LUT3 mapping: priority-cut enumeration (k = 3) + area-flow selection. One LUT3 is one LOP3.
//   4. Two schedules of the same LUT network, each with its own physical register allocation:
//      level order (instructions of a level are independent; the cooperative kernel splits them
//      across the threads of a CTA) and lane order (depth-first, small live set; the lane-parallel
//      kernels run it sequentially per thread).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <queue>

#include "host.hpp"

namespace clg {

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;

// CLG_TIMING=1 prints the wall time of each phase of the pass to stderr
struct Phase {
  const char *name;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit Phase(const char *n) : name(n) {}
  ~Phase() {
    static const bool on = std::getenv("CLG_TIMING") != nullptr;
    if (on) std::fprintf(stderr, "[clg] %-22s %8.3f s\n", name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
};

// ---- and-inverter graph ---------------------------------------------------------------------
// literal = 2 * node + complement; node 0 is constant false (literal 0 = false, 1 = true).
struct Aig {
  std::vector<uint32_t> f0, f1;  // AND fanin literals; inputs: f0 = NONE, f1 = input row
  std::vector<uint32_t> table;   // open-addressing structural hash: node ids, NONE = empty
  size_t used = 0;

  Aig() {
    f0.push_back(NONE), f1.push_back(NONE);
    table.assign(1u << 12, NONE);
  }
  size_t size() const { return f0.size(); }
  bool is_and(uint32_t n) const { return f0[n] != NONE; }
  bool is_input(uint32_t n) const { return n != 0 && f0[n] == NONE; }
  uint32_t input(uint32_t row) {
    f0.push_back(NONE), f1.push_back(row);
    return (uint32_t)(f0.size() - 1) << 1;
  }
  static uint64_t mix(uint32_t a, uint32_t b) {
    uint64_t x = ((uint64_t)a << 32) | b;
    x ^= x >> 33, x *= 0xff51afd7ed558ccdULL, x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ULL, x ^= x >> 33;
    return x;
  }
  void grow() {
    std::vector<uint32_t> old;
    old.swap(table);
    table.assign(old.size() * 2, NONE);
    for (uint32_t n : old)
      if (n != NONE) {
        size_t h = mix(f0[n], f1[n]) & (table.size() - 1);
        while (table[h] != NONE) h = (h + 1) & (table.size() - 1);
        table[h] = n;
      }
  }
  uint32_t and_(uint32_t a, uint32_t b) {
    if (a > b) std::swap(a, b);
    if (a == 0) return 0;
    if (a == 1) return b;
    if (a == b) return a;
    if ((a ^ 1u) == b) return 0;
    size_t h = mix(a, b) & (table.size() - 1);
    while (table[h] != NONE) {
      uint32_t n = table[h];
      if (f0[n] == a && f1[n] == b) return n << 1;
      h = (h + 1) & (table.size() - 1);
    }
    if (f0.size() >= 0x7FFFFFF0u) fail(CLG_E_UNSUPPORTED, "program too large (2^31 graph nodes)");
    f0.push_back(a), f1.push_back(b);
    uint32_t n = (uint32_t)(f0.size() - 1);
    table[h] = n;
    if (++used * 2 > table.size()) grow();
    return n << 1;
  }
};

// ---- truth tables over <= 3 ordered leaves -----------------------------------------------------
// Bit ((v0 << 2) | (v1 << 1) | v2) of a table is the function value; leaf i is LOP3 operand a/b/c.
constexpr uint8_t VAR[3] = {0xF0, 0xCC, 0xAA};

// Re-express table `tt` (over leaves placed at positions 0..2) with leaf i moved to position p[i].
inline uint8_t remap(uint8_t tt, const uint8_t p[3]) {
  uint8_t r = 0;
  for (int m = 0; m < 8; m++) {
    int idx = ((((m >> (2 - p[0])) & 1) << 2) | (((m >> (2 - p[1])) & 1) << 1) | ((m >> (2 - p[2])) & 1));
    r |= (uint8_t)(((tt >> idx) & 1) << m);
  }
  return r;
}

struct RemapTable {
  uint8_t t[27][256];
  RemapTable() {
    for (int q = 0; q < 27; q++) {
      uint8_t p[3] = {(uint8_t)(q / 9), (uint8_t)((q / 3) % 3), (uint8_t)(q % 3)};
      for (int tt = 0; tt < 256; tt++) t[q][tt] = remap((uint8_t)tt, p);
    }
  }
};
const RemapTable &remap_table() {
  static RemapTable T;
  return T;
}

struct Cut {
  uint32_t leaf[3];
  uint8_t n;
  uint8_t tt;
  uint16_t depth;
  float flow;
};

// The mapped network handed to the schedulers.
struct Val {
  uint8_t is_lut;  // 0: input row
  uint8_t tt;
  uint8_t nfan;
  uint32_t fan[3];
  uint32_t row;  // input row for is_lut == 0
};
struct OutRef {
  uint32_t row;
  uint32_t val;  // NONE: constant
  bool inv;      // invert on store / constant value
};

struct Network {
  std::vector<Val> vals;  // topological: every fanin index < own index
  std::vector<OutRef> outs;
  uint32_t n_inputs = 0, n_state = 0, n_outputs = 0;
};

// LUT inputs commute if the truth table is permuted with them. Put every LUT into the operand order with the
// smallest table: functions that differ only by operand order (e.g. the three placements of the odd input of
// NAND(NAND(p,q),y)) collapse onto one LOP3 immediate, which keeps warps of the cooperative kernel on one
// dispatch target.
void canonical_operand_order(Val &v) {
  static const uint8_t P3[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  if (v.nfan < 2) return;
  uint8_t best_tt = v.tt;
  int best = 0;
  for (int q = 1; q < 6; q++) {
    if (v.nfan == 2 && P3[q][2] != 2) continue;
    uint8_t t = remap(v.tt, P3[q]);
    if (t < best_tt) best_tt = t, best = q;
  }
  if (!best) return;
  uint32_t f[3] = {v.fan[0], v.fan[1], v.fan[2]};
  for (int k = 0; k < v.nfan; k++) v.fan[P3[best][k]] = f[k];
  v.tt = best_tt;
}

// ---- LUT3 mapping ------------------------------------------------------------------------------
struct Mapper {
  const Aig &g;
  const std::vector<uint32_t> &out_lits;
  bool fuse;
  int P;  // cuts kept per node

  std::vector<uint8_t> live;
  std::vector<uint32_t> refs;      // AIG fanout counts (reachable part)
  std::vector<float> flow;         // best area flow per node
  std::vector<uint16_t> depth;     // best depth per node
  std::vector<Cut> best;           // chosen cut per AND node
  std::vector<Cut> pool;           // all kept cuts
  std::vector<uint32_t> pool_off;  // per node
  std::vector<uint8_t> pool_cnt;

  Mapper(const Aig &g_, const std::vector<uint32_t> &o, bool fuse_) : g(g_), out_lits(o), fuse(fuse_), P(6) {}

  void mark_live() {
    size_t N = g.size();
    live.assign(N, 0);
    refs.assign(N, 0);
    for (uint32_t l : out_lits) live[l >> 1] = 1, refs[l >> 1]++;
    for (size_t n = N; n-- > 1;)
      if (live[n] && g.is_and((uint32_t)n)) {
        live[g.f0[n] >> 1] = 1, refs[g.f0[n] >> 1]++;
        live[g.f1[n] >> 1] = 1, refs[g.f1[n] >> 1]++;
      }
  }

  static int merge(const Cut &x, const Cut &y, uint32_t out[3]) {
    int i = 0, j = 0, k = 0;
    while (i < x.n || j < y.n) {
      uint32_t v;
      if (j >= y.n || (i < x.n && x.leaf[i] < y.leaf[j])) v = x.leaf[i++];
      else if (i >= x.n || y.leaf[j] < x.leaf[i]) v = y.leaf[j++];
      else v = x.leaf[i], i++, j++;
      if (k == 3) return -1;
      out[k++] = v;
    }
    return k;
  }
  static uint8_t expand(const Cut &c, const uint32_t *leaf, int n) {
    uint8_t p[3] = {0, 0, 0};
    for (int i = 0; i < c.n; i++)
      for (int j = 0; j < n; j++)
        if (leaf[j] == c.leaf[i]) p[i] = (uint8_t)j;
    return remap_table().t[p[0] * 9 + p[1] * 3 + p[2]][c.tt];
  }

  float leaf_w = 0.f;  // cost of a cut = 1 + leaf_w * leaves (0: count LUTs; > 0: also count operand loads, what the cooperative kernels pay per LUT)
  void eval_cut(Cut &c, uint32_t node) const {
    float f = 1.0f + leaf_w * (float)c.n;
    uint16_t d = 0;
    for (int i = 0; i < c.n; i++) f += flow[c.leaf[i]], d = std::max(d, depth[c.leaf[i]]);
    c.flow = f / (float)std::max<uint32_t>(1u, refs[node]);
    c.depth = (uint16_t)std::min<int>(65535, d + 1);
  }
  static bool better(const Cut &a, const Cut &b) {
    if (a.flow != b.flow) return a.flow < b.flow;
    if (a.depth != b.depth) return a.depth < b.depth;
    return a.n < b.n;
  }

  void enumerate() {
    size_t N = g.size();
    flow.assign(N, 0.f), depth.assign(N, 0);
    best.assign(N, Cut{});
    pool_off.assign(N, 0), pool_cnt.assign(N, 0);
    pool.clear();
    std::vector<Cut> cand, sx, sy;
    for (uint32_t n = 1; n < N; n++) {
      if (!live[n] || !g.is_and(n)) continue;
      uint32_t x = g.f0[n] >> 1, y = g.f1[n] >> 1;
      uint8_t nx = (g.f0[n] & 1) ? 0xFF : 0, ny = (g.f1[n] & 1) ? 0xFF : 0;
      auto gather = [&](uint32_t v, std::vector<Cut> &s) {
        s.clear();
        Cut t{};
        t.leaf[0] = v, t.n = 1, t.tt = VAR[0];
        s.push_back(t);
        if (fuse)
          for (int i = 0; i < pool_cnt[v]; i++) s.push_back(pool[pool_off[v] + i]);
      };
      gather(x, sx), gather(y, sy);
      cand.clear();
      for (const Cut &cx : sx)
        for (const Cut &cy : sy) {
          Cut c{};
          int k = merge(cx, cy, c.leaf);
          if (k < 0) continue;
          c.n = (uint8_t)k;
          bool dup = false;
          for (const Cut &o : cand)
            if (o.n == c.n && o.leaf[0] == c.leaf[0] && (c.n < 2 || o.leaf[1] == c.leaf[1]) &&
                (c.n < 3 || o.leaf[2] == c.leaf[2])) {
              dup = true;
              break;
            }
          if (dup) continue;
          c.tt = (uint8_t)((expand(cx, c.leaf, k) ^ nx) & (expand(cy, c.leaf, k) ^ ny));
          eval_cut(c, n);
          cand.push_back(c);
        }
      std::sort(cand.begin(), cand.end(), better);
      if ((int)cand.size() > P) cand.resize(P);
      best[n] = cand[0];
      flow[n] = cand[0].flow, depth[n] = cand[0].depth;
      pool_off[n] = (uint32_t)pool.size(), pool_cnt[n] = (uint8_t)cand.size();
      pool.insert(pool.end(), cand.begin(), cand.end());
    }
  }

  // Re-pick best cuts with fanout estimates taken from the current cover (area recovery).
  std::vector<uint32_t> cover_refs() const {
    size_t N = g.size();
    std::vector<uint32_t> r(N, 0);
    for (uint32_t l : out_lits) r[l >> 1]++;
    for (size_t n = N; n-- > 1;)
      if (r[n] && g.is_and((uint32_t)n))
        for (int i = 0; i < best[n].n; i++) r[best[n].leaf[i]]++;
    return r;
  }
  size_t cover_size() const {
    std::vector<uint32_t> r = cover_refs();
    size_t c = 0;
    for (size_t n = 1; n < g.size(); n++) c += (r[n] && g.is_and((uint32_t)n));
    return c;
  }
  void reselect() {
    std::vector<uint32_t> cr = cover_refs();
    size_t N = g.size();
    for (size_t n = 1; n < N; n++) refs[n] = std::max<uint32_t>(1u, (uint32_t)((cr[n] * 2 + refs[n] + 1) / 3));
    for (uint32_t n = 1; n < N; n++) {
      if (!live[n] || !g.is_and(n)) continue;
      Cut *c = &pool[pool_off[n]];
      int bi = 0;
      for (int i = 0; i < pool_cnt[n]; i++) {
        eval_cut(c[i], n);
        if (better(c[i], c[bi])) bi = i;
      }
      best[n] = c[bi];
      flow[n] = c[bi].flow, depth[n] = c[bi].depth;
    }
  }

  // Exact-area local refinement: for every node in the cover try each of its cuts and keep the
  // one that minimises the number of LUTs that would become (or stay) referenced.
  struct Exact {
    Mapper &m;
    std::vector<uint32_t> r;
    explicit Exact(Mapper &m_) : m(m_), r(m_.cover_refs()) {}
    uint32_t cost(uint32_t n) const { return 256u + (uint32_t)(256.f * m.leaf_w) * m.best[n].n; }
    uint32_t deref(uint32_t n) {
      uint32_t a = cost(n);
      for (int i = 0; i < m.best[n].n; i++) {
        uint32_t l = m.best[n].leaf[i];
        if (--r[l] == 0 && m.g.is_and(l)) a += deref(l);
      }
      return a;
    }
    uint32_t ref(uint32_t n) {
      uint32_t a = cost(n);
      for (int i = 0; i < m.best[n].n; i++) {
        uint32_t l = m.best[n].leaf[i];
        if (r[l]++ == 0 && m.g.is_and(l)) a += ref(l);
      }
      return a;
    }
    void run() {
      size_t N = m.g.size();
      for (uint32_t n = 1; n < N; n++) {
        if (!r[n] || !m.g.is_and(n)) continue;
        deref(n);
        Cut keep = m.best[n];
        uint32_t best_area = NONE;
        Cut best_cut = keep;
        Cut *c = &m.pool[m.pool_off[n]];
        for (int i = 0; i < m.pool_cnt[n]; i++) {
          m.best[n] = c[i];
          uint32_t a = ref(n);
          deref(n);
          if (a < best_area || (a == best_area && c[i].depth < best_cut.depth)) best_area = a, best_cut = c[i];
        }
        m.best[n] = best_cut;
        ref(n);
      }
    }
  };

  void run() {
    if (const char *env = std::getenv("CLG_MAP_LEAF_WEIGHT")) leaf_w = (float)std::atof(env);  // (experiment knob)
    {
      Phase ph("map: cut enumeration");
      mark_live();
      enumerate();
    }
    if (!fuse) return;
    {
      Phase ph("map: area flow");
      for (int it = 0; it < 2; it++) reselect();
    }
    // deep recursion in Exact is bounded by cone size, not graph depth, but guard huge graphs
    Phase ph("map: exact area");
    if (g.size() < (1u << 24))
      for (int it = 0; it < 2; it++) Exact(*this).run();
  }
};

// ---- schedules + register allocation -----------------------------------------------------------
struct Sched {
  std::vector<FlatInstr> instr;
  std::vector<uint32_t> levels;
  uint32_t n_slots = 0, max_width = 0;
  bool ok = true;
};

struct SlotPool {
  std::priority_queue<uint32_t, std::vector<uint32_t>, std::greater<uint32_t>> free_;
  uint32_t high = 0;
  uint32_t get() {
    if (!free_.empty()) {
      uint32_t s = free_.top();
      free_.pop();
      return s;
    }
    return high++;
  }
  void put(uint32_t s) { free_.push(s); }
};

FlatInstr make_lut(const Val &v) {
  FlatInstr in{};
  in.kind = K_LUT, in.tt = v.tt;
  in.flags = (uint16_t)(v.nfan << NFAN_SHIFT);
  return in;
}

// Lane order: depth-first from the outputs (taller operand first), inputs loaded at first use,
// outputs stored as soon as they exist. Slots are recycled the moment their last reader is done.
Sched schedule_lane(const Network &net) {
  const size_t V = net.vals.size();
  std::vector<uint32_t> height(V, 0);
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut) {
      uint32_t h = 0;
      for (int i = 0; i < net.vals[v].nfan; i++) h = std::max(h, height[net.vals[v].fan[i]]);
      height[v] = h + 1;
    }
  std::vector<std::vector<uint32_t>> outs_of(V);
  std::vector<uint32_t> state_in_val(net.n_state, NONE);  // value that loads state row k
  for (size_t v = 0; v < V; v++)
    if (!net.vals[v].is_lut && net.vals[v].row >= net.n_inputs) state_in_val[net.vals[v].row - net.n_inputs] = (uint32_t)v;
  for (size_t o = 0; o < net.outs.size(); o++)
    if (net.outs[o].val != NONE) outs_of[net.outs[o].val].push_back((uint32_t)o);
  std::vector<std::vector<uint32_t>> users(V);  // LUTs reading each carried-state load
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut)
      for (int i = 0; i < net.vals[v].nfan; i++) {
        const uint32_t f = net.vals[v].fan[i];
        if (!net.vals[f].is_lut && net.vals[f].row >= net.n_inputs && (users[f].empty() || users[f].back() != (uint32_t)v)) users[f].push_back((uint32_t)v);
      }

  // order: positive = value id (LDIN or LUT), stores as (V + out index). Depth-first from the outputs, taller
  // operand first. (A greedy "least live-set growth" list scheduler was tried and was 3-13x worse on the
  // multipliers: it opens every partial product before closing any adder row.)
  std::vector<uint32_t> order;
  auto build_dfs = [&]() {
  order.clear();
  order.reserve(V + net.outs.size());
  std::vector<uint8_t> done(V, 0), stored(net.outs.size(), 0);
  // emission is driven by a work list (no recursion: a chain of 10^6 latches is 10^6 nested forced loads)
  std::vector<uint32_t> work;  // value id, or V + out index for a store
  auto emit_value = [&](uint32_t v) {
    done[v] = 1;
    order.push_back(v);
    for (size_t k = outs_of[v].size(); k-- > 0;) work.push_back((uint32_t)V + outs_of[v][k]);
  };
  auto drain = [&]() {
    while (!work.empty()) {
      const uint32_t x = work.back();
      work.pop_back();
      if (x < V) {  // a reader of a forced load that was ready when it was queued
        if (!done[x]) emit_value(x);
        continue;
      }
      const uint32_t o = x - (uint32_t)V;
      if (stored[o]) continue;
      const uint32_t row = net.outs[o].row;
      uint32_t forced = NONE;
      if (row >= net.n_outputs) {  // in-place state row: its old value must have been read first
        const uint32_t sv = state_in_val[row - net.n_outputs];
        if (sv != NONE && !done[sv]) done[sv] = 1, order.push_back(sv), forced = sv;
      }
      stored[o] = 1;
      order.push_back((uint32_t)V + o);
      if (forced != NONE) {
        // The old value was loaded only to get it out of the way of this store; whoever reads it may be far down
        // the depth-first order, and it would sit in a register until then (a chain of n latches kept 2n values
        // live). Run the readers that are ready right now -- their own stores may free further rows the same way.
        for (size_t k = users[forced].size(); k-- > 0;) {
          const uint32_t c = users[forced][k];
          if (done[c]) continue;
          bool ready = true;
          for (int i = 0; i < net.vals[c].nfan; i++) ready &= done[net.vals[c].fan[i]] != 0;
          if (ready) work.push_back(c);
        }
        for (size_t k = outs_of[forced].size(); k-- > 0;) work.push_back((uint32_t)V + outs_of[forced][k]);
      }
    }
  };
  struct Frame {
    uint32_t v;
    uint8_t next;
    uint32_t kid[3];
    uint8_t nk;
  };
  std::vector<Frame> stack;
  auto visit = [&](uint32_t root) {
    if (done[root]) return;
    auto open = [&](uint32_t v) {
      Frame f{v, 0, {0, 0, 0}, 0};
      const Val &val = net.vals[v];
      if (val.is_lut) {
        f.nk = val.nfan;
        for (int i = 0; i < val.nfan; i++) f.kid[i] = val.fan[i];
        std::stable_sort(f.kid, f.kid + f.nk, [&](uint32_t a, uint32_t b) { return height[a] > height[b]; });
      }
      stack.push_back(f);
    };
    open(root);
    while (!stack.empty()) {
      Frame &f = stack.back();
      if (done[f.v]) {
        stack.pop_back();
        continue;
      }
      if (f.next < f.nk) {
        uint32_t k = f.kid[f.next++];
        if (!done[k]) open(k);
        continue;
      }
      uint32_t v = f.v;
      stack.pop_back();
      emit_value(v);
      drain();
    }
  };
  for (size_t o = 0; o < net.outs.size(); o++) {
    if (net.outs[o].val != NONE) visit(net.outs[o].val);
    work.push_back((uint32_t)V + (uint32_t)o);  // (a constant stored to an in-place state row also waits for that row's load)
    drain();
  }
  };

  build_dfs();

  auto allocate = [&](const std::vector<uint32_t> &seq) {
  // last use positions
  std::vector<uint32_t> last(V, 0);
  for (size_t i = 0; i < seq.size(); i++) {
    uint32_t x = seq[i];
    if (x >= V) {
      const OutRef &o = net.outs[x - V];
      if (o.val != NONE) last[o.val] = (uint32_t)i;
    } else if (net.vals[x].is_lut)
      for (int k = 0; k < net.vals[x].nfan; k++) last[net.vals[x].fan[k]] = (uint32_t)i;
  }
  Sched s;
  SlotPool pool;
  std::vector<uint32_t> slot(V, NONE);
  for (size_t i = 0; i < seq.size(); i++) {
    uint32_t x = seq[i];
    FlatInstr in{};
    if (x >= V) {
      const OutRef &o = net.outs[x - V];
      in.row = o.row, in.flags = o.inv ? 1 : 0;
      if (o.val == NONE) in.kind = K_STCONST;
      else {
        in.kind = K_STOUT, in.a = (uint16_t)slot[o.val];
        if (last[o.val] == i) pool.put(slot[o.val]);
      }
    } else {
      const Val &v = net.vals[x];
      if (v.is_lut) {
        in = make_lut(v);
        uint32_t sl[3] = {0, 0, 0};
        for (int k = 0; k < v.nfan; k++) sl[k] = slot[v.fan[k]];
        for (int k = v.nfan; k < 3; k++) sl[k] = sl[0];  // unused operands alias operand a
        in.a = (uint16_t)sl[0], in.b = (uint16_t)sl[1], in.c = (uint16_t)sl[2];
        for (int k = 0; k < v.nfan; k++) {
          uint32_t fv = v.fan[k];
          bool seen = false;
          for (int j = 0; j < k; j++) seen |= (v.fan[j] == fv);
          if (!seen && last[fv] == i) pool.put(slot[fv]);
        }
      } else {
        in.kind = K_LDIN, in.row = v.row;
      }
      slot[x] = pool.get();
      if (slot[x] > 0xFFFF) {
        s.ok = false;
        return s;
      }
      in.dst = (uint16_t)slot[x];
    }
    s.instr.push_back(in);
  }
  s.n_slots = pool.high;
  s.levels = {0, (uint32_t)s.instr.size()};
  s.max_width = (uint32_t)s.instr.size();
  return s;
  };

  Sched s = allocate(order);
  // Second candidate: the order the program was written in. Values are numbered in source order (the SSA pass walks
  // the ops front to back), inputs are loaded at first use, outputs stored as soon as they exist. A demand-driven
  // depth-first walk computes exactly the cone of each output, which for a row-by-row design (an array multiplier)
  // is a diagonal through every row -- two live values per row, 315 for the 64x64 multiplier -- while its author's
  // order finishes one row before the next. Keep whichever needs fewer registers.
  {
    std::vector<uint32_t> src;
    src.reserve(order.size());
    std::vector<uint8_t> done(V, 0), stored(net.outs.size(), 0);
    auto load = [&](uint32_t v) {
      if (!done[v]) done[v] = 1, src.push_back(v);
    };
    auto store = [&](uint32_t o) {
      if (stored[o]) return;
      const uint32_t row = net.outs[o].row;
      if (row >= net.n_outputs && state_in_val[row - net.n_outputs] != NONE) load(state_in_val[row - net.n_outputs]);
      stored[o] = 1, src.push_back((uint32_t)V + o);
    };
    for (uint32_t v = 0; v < V; v++) {
      if (!net.vals[v].is_lut) continue;
      for (int i = 0; i < net.vals[v].nfan; i++)
        if (!net.vals[net.vals[v].fan[i]].is_lut) load(net.vals[v].fan[i]);
      done[v] = 1, src.push_back(v);
      for (uint32_t o : outs_of[v]) store(o);
    }
    for (uint32_t o = 0; o < net.outs.size(); o++) {  // constants, and inputs stored straight to outputs
      if (net.outs[o].val != NONE) load(net.outs[o].val);
      store(o);
    }
    for (uint32_t v = 0; v < V; v++) load(v);  // (inputs nothing reads still get their slot, as in the other order)
    Sched alt = allocate(src);
    if (alt.ok && (!s.ok || alt.n_slots < s.n_slots)) s = std::move(alt), order.swap(src);
  }
  // Third candidate: a forward list schedule that always runs the ready instruction with the best register balance
  // (values it kills minus the one it creates), and among equals the one whose reader is closest to being ready, then
  // source order. On a row-by-row design this walks along the rows instead of across them.
  // (inputs are offered to the scheduler like any other instruction, or loaded on demand right before their first
  // reader -- the multipliers: 132 live values instead of 222, a thin DAG: 86 instead of 77)
  for (const bool on_demand : {false, true}) {
    if (std::getenv("CLG_LANE_NO_LIST")) break;
    // (no order will bring a live set of thousands down to the few hundred the lane modes can use: do not spend the
    // second on it -- the 1M-gate DAG: 6 622 live values in the orders above, 5 252 in this one)
    if (s.ok && s.n_slots > 3000) break;
    // The on-demand order is the tightest (it finishes a multiplier row before it starts the next) and for that very
    // reason the slowest to run when both fit the register file: one dependent carry chain per warp, no second row to
    // overlap with (mul64 x 12: 7.9 ms against 7.2 ms as one kernel per instance; 5.57 against 5.46-5.56 ms in the
    // second chain, where it hands 128 registers over instead of 203 and runs three warps per scheduler: no gain
    // either). It is only tried when everything else would spill.
    if (on_demand && s.ok && s.n_slots <= 230 && !std::getenv("CLG_LANE_ON_DEMAND")) break;  // (knob: measure the tight order anyway)
    std::vector<uint32_t> coff(V + 1, 0), cons;  // readers (LUTs) of each value
    auto each_operand = [&](uint32_t v, auto &&fn) {
      const Val &x = net.vals[v];
      for (int i = 0; i < x.nfan; i++) {
        bool seen = false;
        for (int j = 0; j < i; j++) seen |= x.fan[j] == x.fan[i];
        if (!seen) fn(x.fan[i]);
      }
    };
    for (uint32_t v = 0; v < V; v++)
      if (net.vals[v].is_lut) each_operand(v, [&](uint32_t f) { coff[f + 1]++; });
    for (uint32_t v = 0; v < V; v++) coff[v + 1] += coff[v];
    cons.resize(coff[V]);
    {
      std::vector<uint32_t> fill(coff.begin(), coff.end() - 1);
      for (uint32_t v = 0; v < V; v++)
        if (net.vals[v].is_lut) each_operand(v, [&](uint32_t f) { cons[fill[f]++] = v; });
    }
    std::vector<uint32_t> uses_left(V), need(V, 0), have(V, 0);  // readers not run yet (+ stores); operands: wanted / done
    for (uint32_t v = 0; v < V; v++) {
      uses_left[v] = coff[v + 1] - coff[v];
      if (net.vals[v].is_lut) each_operand(v, [&](uint32_t f) { need[v] += !on_demand || net.vals[f].is_lut; });
    }
    std::vector<uint8_t> done(V, 0), stored(net.outs.size(), 0), queued(V, 0);
    std::vector<uint64_t> key_of(V, ~0ull);
    std::priority_queue<std::pair<uint64_t, uint32_t>, std::vector<std::pair<uint64_t, uint32_t>>, std::greater<>> heap;
    auto key = [&](uint32_t v) {
      int delta = 1;  // the value it creates
      uint32_t close = 0;
      if (net.vals[v].is_lut)
        each_operand(v, [&](uint32_t f) {
          delta -= uses_left[f] == 1 && outs_of[f].empty();
          delta += on_demand && !net.vals[f].is_lut && !done[f];  // an input it would have to load first
        });
      for (uint32_t k = coff[v]; k < coff[v + 1]; k++) {
        const uint32_t c = cons[k];
        close = std::max(close, need[c] ? (have[c] + 1) * 4 / need[c] : 0u);
      }
      return (uint64_t)(delta + 4) << 40 | (uint64_t)(7 - std::min(close, 7u)) << 36 | v;
    };
    auto offer = [&](uint32_t v) {
      if (done[v]) return;
      const uint64_t k = key(v);
      if (queued[v] && k >= key_of[v]) return;
      queued[v] = 1, key_of[v] = k, heap.push({k, v});
    };
    std::vector<uint32_t> lst;
    lst.reserve(order.size());
    std::function<void(uint32_t)> run_value;
    auto store = [&](uint32_t o) {
      if (stored[o]) return;
      const uint32_t row = net.outs[o].row;
      if (row >= net.n_outputs) {
        const uint32_t sv = state_in_val[row - net.n_outputs];
        if (sv != NONE && !done[sv]) run_value(sv);
      }
      stored[o] = 1, lst.push_back((uint32_t)V + o);
    };
    run_value = [&](uint32_t v) {
      if (net.vals[v].is_lut)
        each_operand(v, [&](uint32_t f) {
          if (on_demand && !net.vals[f].is_lut && !done[f]) run_value(f);  // load at first use
        });
      done[v] = 1, lst.push_back(v);
      if (on_demand && !net.vals[v].is_lut) {
        for (uint32_t o : outs_of[v]) store(o);
        // readers that were waiting only for balance reasons see one load less to pay for
        for (uint32_t k = coff[v]; k < coff[v + 1]; k++)
          if (!done[cons[k]] && have[cons[k]] == need[cons[k]]) offer(cons[k]);
        return;
      }
      if (net.vals[v].is_lut)
        each_operand(v, [&](uint32_t f) {
          if (--uses_left[f] == 1)  // its last reader now kills it: that reader's balance improved
            for (uint32_t k = coff[f]; k < coff[f + 1]; k++)
              if (!done[cons[k]] && have[cons[k]] == need[cons[k]]) offer(cons[k]);
        });
      for (uint32_t o : outs_of[v]) store(o);
      for (uint32_t k = coff[v]; k < coff[v + 1]; k++) {
        const uint32_t c = cons[k];
        if (++have[c] == need[c]) offer(c);
        else  // the other operands of this reader are now closer to enabling it
          each_operand(c, [&](uint32_t f) {
            if (!done[f] && (net.vals[f].is_lut ? have[f] == need[f] : !on_demand)) offer(f);
          });
      }
    };
    // one connected component after the other (flattened instances: finish one before opening the next)
    std::vector<uint32_t> comp(V);
    for (uint32_t v = 0; v < V; v++) comp[v] = v;
    auto find = [&](uint32_t x) {
      while (comp[x] != x) x = comp[x] = comp[comp[x]];
      return x;
    };
    for (uint32_t v = 0; v < V; v++)
      if (net.vals[v].is_lut) each_operand(v, [&](uint32_t f) { comp[find(v)] = find(f); });
    std::vector<uint32_t> seeds;
    for (uint32_t v = 0; v < V; v++)
      if (on_demand ? net.vals[v].is_lut && need[v] == 0 : (!net.vals[v].is_lut || need[v] == 0) && uses_left[v] + outs_of[v].size() > 0) seeds.push_back(v);
    std::vector<uint32_t> first_of(V, NONE);  // component root -> its lowest value id
    for (uint32_t v = 0; v < V; v++)
      if (first_of[find(v)] == NONE) first_of[find(v)] = v;
    std::stable_sort(seeds.begin(), seeds.end(), [&](uint32_t a, uint32_t b) { return first_of[find(a)] < first_of[find(b)]; });
    for (size_t i = 0; i < seeds.size();) {
      const uint32_t c = find(seeds[i]);
      for (; i < seeds.size() && find(seeds[i]) == c; i++) offer(seeds[i]);
      while (!heap.empty()) {
        const auto [k, v] = heap.top();
        heap.pop();
        if (done[v] || k != key_of[v]) continue;
        run_value(v);
      }
    }
    for (uint32_t o = 0; o < net.outs.size(); o++) {
      if (net.outs[o].val != NONE && !done[net.outs[o].val]) run_value(net.outs[o].val);
      store(o);
    }
    bool complete = true;
    for (uint32_t v = 0; v < V; v++) {
      if (!done[v] && net.vals[v].is_lut) complete = false;
      if (!done[v] && !net.vals[v].is_lut) done[v] = 1, lst.push_back(v);
    }
    if (complete) {
      Sched alt = allocate(lst);
      if (std::getenv("CLG_TIMING")) std::fprintf(stderr, "[clg]   lane: list schedule %u slots, best other %u\n", alt.n_slots, s.n_slots);
      if (alt.ok && (!s.ok || alt.n_slots < s.n_slots)) s = std::move(alt), order.swap(lst);
    }
  }
  if (!s.ok || net.n_state == 0) return s;
  // Carried-state rows are rewritten in place, so their loads are ordinary (coherent) loads that neither the
  // assembler nor the hardware may move above an earlier store: "load at first use" leaves one load in flight per
  // thread, and a latch chain ran at one memory latency per latch. Hoist each state load up to LOOKAHEAD positions,
  // at most `budget` of them in flight -- as many as the registers left over at this program's K allow.
  const uint32_t k0 = s.n_slots * 4 <= 200 ? 4 : s.n_slots * 2 <= 200 ? 2 : 1;
  const uint32_t budget = 200 / k0 > s.n_slots ? std::min<uint32_t>(24, 200 / k0 - s.n_slots) : 0;
  if (budget < 4) return s;
  constexpr size_t LOOKAHEAD = 512;
  auto is_state_load = [&](uint32_t x) { return x < V && !net.vals[x].is_lut && net.vals[x].row >= net.n_inputs; };
  std::vector<uint32_t> hoisted;
  hoisted.reserve(order.size());
  std::vector<uint8_t> pulled(V, 0);
  size_t scan = 0;
  uint32_t in_flight = 0;
  for (size_t i = 0; i < order.size(); i++) {
    scan = std::max(scan, i + 1);
    while (in_flight < budget && scan < order.size() && scan <= i + LOOKAHEAD) {
      const uint32_t y = order[scan++];
      if (is_state_load(y)) hoisted.push_back(y), pulled[y] = 1, in_flight++;
    }
    const uint32_t x = order[i];
    if (x < V && pulled[x]) {  // its original position: from here on it counts as an ordinary live value
      in_flight--;
      continue;
    }
    hoisted.push_back(x);
  }
  Sched h = allocate(hoisted);
  return h.ok ? h : s;
}

// Level order. LUT levels are as late as the consumers allow (short live ranges), loads sit one
// level before their first reader, stores one level after their producer. A slot freed by its
// last reader in level L is handed out again from level L + 1 on: within a level other threads may
// still be reading it.
Sched schedule_level(const Network &net, uint32_t BANKS = 8, bool estimate_only = false) {
  const size_t V = net.vals.size(), O = net.outs.size();
  std::vector<uint32_t> lev(V, 0);
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut) {
      uint32_t h = 0;
      for (int i = 0; i < net.vals[v].nfan; i++) h = std::max(h, lev[net.vals[v].fan[i]]);
      lev[v] = h + 1;  // ASAP; inputs count as level 0
    }
  std::vector<uint32_t> ub(V, NONE);
  for (size_t v = V; v-- > 0;) {
    if (net.vals[v].is_lut) {
      if (ub[v] != NONE) lev[v] = ub[v];
      for (int i = 0; i < net.vals[v].nfan; i++) {
        uint32_t f = net.vals[v].fan[i];
        ub[f] = std::min(ub[f], lev[v] - 1);
      }
    }
  }
  // loads: a level that contains a global load cannot end before the data is back (~1 us), so loads are not
  // sprinkled one level before their first reader: a stream with few input rows loads them all at level 0, a
  // stream with many batches them every 16 levels (the latency is paid once per batch, the extra live registers
  // are bounded by 16 levels' worth of inputs).
  size_t n_loads = 0;
  for (size_t v = 0; v < V; v++) n_loads += !net.vals[v].is_lut;
  for (size_t v = 0; v < V; v++)
    if (!net.vals[v].is_lut) {
      const uint32_t alap = (ub[v] == NONE) ? 0 : ub[v];
      lev[v] = n_loads <= 2048 ? 0 : alap / 16 * 16;
    }
  std::vector<uint32_t> state_in_lev(net.n_state, NONE);
  for (size_t v = 0; v < V; v++)
    if (!net.vals[v].is_lut && net.vals[v].row >= net.n_inputs) state_in_lev[net.vals[v].row - net.n_inputs] = lev[v];
  std::vector<uint32_t> olev(O, 0);
  uint32_t n_levels = 1;
  for (size_t o = 0; o < O; o++) {
    const OutRef &r = net.outs[o];
    uint32_t l = (r.val == NONE) ? 0 : lev[r.val] + 1;
    if (r.row >= net.n_outputs) {
      uint32_t sl = state_in_lev[r.row - net.n_outputs];
      if (sl != NONE) l = std::max(l, sl + 1);
    }
    olev[o] = l;
    n_levels = std::max(n_levels, l + 1);
  }
  for (size_t v = 0; v < V; v++) n_levels = std::max(n_levels, lev[v] + 1);

  std::vector<uint32_t> last(V, 0);
  for (size_t v = 0; v < V; v++) {
    last[v] = std::max(last[v], lev[v]);
    if (net.vals[v].is_lut)
      for (int i = 0; i < net.vals[v].nfan; i++) last[net.vals[v].fan[i]] = std::max(last[net.vals[v].fan[i]], lev[v]);
  }
  for (size_t o = 0; o < O; o++)
    if (net.outs[o].val != NONE) last[net.outs[o].val] = std::max(last[net.outs[o].val], olev[o]);

  if (estimate_only) {  // peak number of values alive in one level: what the register stack will need, within a few %
    std::vector<int64_t> diff(n_levels + 2, 0);
    for (size_t v = 0; v < V; v++) diff[lev[v]]++, diff[last[v] + 1]--;
    int64_t live = 0, peak = 0;
    for (uint32_t l = 0; l < n_levels; l++) live += diff[l], peak = std::max(peak, live);
    Sched est;
    est.n_slots = (uint32_t)peak;
    return est;
  }
  // bucket by level
  std::vector<uint32_t> cnt(n_levels + 1, 0);
  for (size_t v = 0; v < V; v++) cnt[lev[v] + 1]++;
  for (size_t o = 0; o < O; o++) cnt[olev[o] + 1]++;
  for (uint32_t l = 0; l < n_levels; l++) cnt[l + 1] += cnt[l];
  std::vector<uint32_t> items(V + O), fill(cnt.begin(), cnt.end() - 1);
  for (size_t v = 0; v < V; v++) items[fill[lev[v]]++] = (uint32_t)v;
  for (size_t o = 0; o < O; o++) items[fill[olev[o]]++] = (uint32_t)(V + o);

  // ---- emission: bank-aware order, operand placement and slot choice -------------------------------------
  // The cooperative kernel keeps the stack as [slot][K words]. At K = 4 a slot is 16 bytes, 8 consecutive slots
  // span the 32 banks, and an LDS.128/STS.128 is served one quarter-warp (8 lanes = 8 consecutive instructions of
  // the level) at a time. A quarter is conflict-free iff its 8 addresses fall in 8 different 16-byte bank
  // groups, i.e. slot % 8 all different -- separately for operand a, b, c and the destination, since each is
  // its own LDS/STS. (K = 2: half-warps and slot % 16; K = 1: whole warps and slot % 32 -- BANKS below. The
  // caller picks BANKS for the K the register stack will fit shared memory with.) Three degrees of freedom are used: which 8 instructions share a quarter (greedy packing
  // from a window of candidates), which operand sits in which position (LUT inputs commute if the truth table
  // is permuted with them), and which free slot a result takes (one bank group per quarter member).
  constexpr uint32_t MAX_BANKS = 32;
  Sched s;
  s.instr.resize(V + O);
  s.levels.assign(cnt.begin(), cnt.end());
  struct BankPool {
    uint32_t BANKS;
    std::vector<uint32_t> free_[MAX_BANKS];
    uint32_t fresh[MAX_BANKS];
    uint32_t high = 0;
    explicit BankPool(uint32_t nb) : BANKS(nb) {
      for (uint32_t b = 0; b < BANKS; b++) fresh[b] = b;
    }
    bool has_free(uint32_t b) const { return !free_[b].empty(); }
    uint32_t get(uint32_t b) {
      uint32_t sl;
      if (!free_[b].empty()) sl = free_[b].back(), free_[b].pop_back();
      else sl = fresh[b], fresh[b] += BANKS;
      high = std::max(high, sl + 1);
      return sl;
    }
    void put(uint32_t sl) { free_[sl % BANKS].push_back(sl); }
  } pool(BANKS);
  std::vector<uint32_t> slot(V, NONE);
  std::vector<std::vector<uint32_t>> release(n_levels + 1);
  auto klass = [&](uint32_t x) -> uint32_t {  // widest LUTs first (the kernel skips loads a LUT does not need)
    if (x >= V) return net.outs[x - V].val == NONE ? 5u : 4u;
    return net.vals[x].is_lut ? (uint32_t)(3 - net.vals[x].nfan) : 3u;
  };
  auto key = [&](uint32_t x) -> uint32_t {
    uint32_t sub = x >= V ? (net.outs[x - V].inv ? 1u : 0u) : net.vals[x].is_lut ? net.vals[x].tt : 0u;
    return klass(x) << 8 | sub;
  };
  static const uint8_t PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  struct Placed {
    uint32_t x;
    uint8_t perm;  // operand k of the value goes to position PERM[perm][k]
  };
  struct CandInfo {
    uint8_t nops, perms, bank[3];
  };
  std::vector<uint32_t> fanout(V, 0);  // operand reads of each value (LUT operands + stores)
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut)
      for (int k = 0; k < net.vals[v].nfan; k++) fanout[net.vals[v].fan[k]]++;
  for (const OutRef &o : net.outs)
    if (o.val != NONE) fanout[o.val]++;
  uint64_t bank_weight[MAX_BANKS] = {};
  // Three phases. (1) Level by level: which instructions share a group, operand placement, and a first choice of the
  // bank group of every result. (2) Local search over the bank choices with every reader known. (3) Slots.
  std::vector<uint8_t> bankv(V, 0);   // bank group of each value
  std::vector<Placed> seq(V + O);     // the final order: level l occupies [cnt[l], cnt[l + 1])
  std::vector<Placed> order;
  std::vector<uint32_t> nxt;
  std::vector<CandInfo> cinfo;
  for (uint32_t l = 0; l < n_levels; l++) {
    const uint32_t b = cnt[l], e = cnt[l + 1], n = e - b;
    s.max_width = std::max(s.max_width, n);
    std::stable_sort(items.begin() + b, items.begin() + e, [&](uint32_t x, uint32_t y) { return key(x) < key(y); });
    // -- pack quarters
    order.clear();
    // per candidate: operand count, operand bank groups, and which operand orders leave its table unchanged
    cinfo.resize(n);
    for (uint32_t c = 0; c < n; c++) {
      const uint32_t x = items[b + c];
      CandInfo &ci = cinfo[c];
      ci.perms = 1;
      if (x >= V) {
        ci.nops = net.outs[x - V].val == NONE ? 0 : 1;
        if (ci.nops) ci.bank[0] = bankv[net.outs[x - V].val];
      } else if (net.vals[x].is_lut) {
        const Val &v = net.vals[x];
        ci.nops = v.nfan;
        for (int k = 0; k < v.nfan; k++) ci.bank[k] = bankv[v.fan[k]];
        if (v.nfan >= 2)
          for (int q = 1; q < 6; q++) {
            bool fits = true;
            for (int k = 0; k < v.nfan; k++) fits &= PERM[q][k] < v.nfan;
            if (fits && remap(v.tt, PERM[q]) == v.tt) ci.perms |= (uint8_t)(1u << q);
          }
      } else {
        ci.nops = 0;
      }
    }
    // how far a group looks for members that fit without a conflict; very wide levels (flattened instances: tens of
    // thousands of similar instructions) have plenty of choice close by, and a long look-ahead there made the pass
    // quadratic
    const uint32_t WINDOW = n > 16384 ? 256 : n > 8192 ? 512 : 2048;
    // (the candidates not placed yet form a linked list in their sorted order: a window scan steps over the ones
    // already taken in O(1) instead of re-reading them for every group)
    nxt.resize(n + 1);
    for (uint32_t c = 0; c <= n; c++) nxt[c] = c + 1;
    uint32_t head = 0, placed = 0;
    auto unlink = [&](uint32_t prev, uint32_t c) {
      if (prev == NONE) head = nxt[c];
      else nxt[prev] = nxt[c];
    };
    while (placed < n) {
      uint32_t mask[3] = {0, 0, 0};
      uint32_t members = 0;
      uint32_t scanned = 0;
      // a group is drawn from ONE class (operand count + truth table) while that class lasts: the row kernel
      // (clg_coop_rows) runs a warp's 32 instructions with one uniform LOP3 dispatch per distinct table, so
      // groups that mix tables cost it an extra pass; the classes are contiguous in the sorted order
      const uint32_t head_key = head < n ? key(items[b + head]) : 0;
      for (uint32_t c = head, prev = NONE; c < n && members < BANKS && scanned < WINDOW;) {
        if (key(items[b + c]) != head_key && !std::getenv("CLG_MIX_CLASSES")) break;
        scanned++;
        const CandInfo &ci = cinfo[c];
        int pick = -1;
        for (int q = 0; q < 6 && pick < 0; q++) {
          if (!(ci.perms >> q & 1)) continue;
          bool ok = true;
          for (int k = 0; k < ci.nops && ok; k++) ok = !(mask[PERM[q][k]] >> ci.bank[k] & 1u);
          if (ok) pick = q;
        }
        if (pick < 0) {
          prev = c, c = nxt[c];
          continue;
        }
        for (int k = 0; k < ci.nops; k++) mask[PERM[pick][k]] |= 1u << ci.bank[k];
        members++, placed++;
        order.push_back({items[b + c], (uint8_t)pick});
        unlink(prev, c);
        c = nxt[c];
      }
      // nothing else in the window fits without a conflict (the tail of a class): fill the group with the
      // candidates that collide least, so that the replays spread over banks instead of piling onto one
      if (members < BANKS && placed < n) {
        uint8_t hits[3][MAX_BANKS] = {};
        for (uint32_t k = order.size() - members; k < order.size(); k++) {
          const uint32_t x = order[k].x;
          const uint8_t *pm = PERM[order[k].perm];
          const int no = x >= V ? (net.outs[x - V].val == NONE ? 0 : 1) : net.vals[x].is_lut ? net.vals[x].nfan : 0;
          for (int j = 0; j < no; j++) hits[pm[j]][x >= V ? bankv[net.outs[x - V].val] : bankv[net.vals[x].fan[j]]]++;
        }
        while (members < BANKS && placed < n) {
          uint32_t best_c = NONE, best_prev = NONE, best_cost = NONE, seen = 0;
          int best_q = 0;
          for (uint32_t c = head, prev = NONE; c < n && seen < 64; prev = c, c = nxt[c]) {
            seen++;
            const CandInfo &ci = cinfo[c];
            for (int q = 0; q < 6; q++) {
              if (!(ci.perms >> q & 1)) continue;
              uint32_t cost = 0;
              for (int k = 0; k < ci.nops; k++) cost += hits[PERM[q][k]][ci.bank[k]];
              if (cost < best_cost) best_cost = cost, best_c = c, best_prev = prev, best_q = q;
            }
            if (best_cost == 0) break;
          }
          const CandInfo &ci = cinfo[best_c];
          for (int k = 0; k < ci.nops; k++) hits[PERM[best_q][k]][ci.bank[k]]++;
          members++, placed++;
          order.push_back({items[b + best_c], (uint8_t)best_q});
          unlink(best_prev, best_c);
        }
      }
    }
    // -- first choice of destination bank groups, quarter by quarter
    // The members of a group take distinct bank groups (conflict-free STS). WHICH member takes which is chosen so
    // that the read traffic per bank group stays level: values are handed out most-read first to the bank group
    // that has accumulated the least fan-out so far. (With a round-robin choice the readers of a level see up to
    // +/-25 % imbalance between bank groups, which no packing of the readers can undo.)
    for (uint32_t i = 0; i < n; i += BANKS) {
      uint32_t idx[MAX_BANKS], m = 0, dmask = 0;
      for (uint32_t k = i; k < std::min(n, i + BANKS); k++)
        if (order[k].x < V) idx[m++] = k;
      std::stable_sort(idx, idx + m, [&](uint32_t p, uint32_t q) { return fanout[order[p].x] > fanout[order[q].x]; });
      for (uint32_t j = 0; j < m; j++) {
        uint32_t bank = NONE;
        for (uint32_t cand = 0; cand < BANKS; cand++) {
          if (dmask >> cand & 1u) continue;
          if (bank == NONE || bank_weight[cand] < bank_weight[bank]) bank = cand;
        }
        dmask |= 1u << bank;
        bank_weight[bank] += fanout[order[idx[j]].x];
        bankv[order[idx[j]].x] = (uint8_t)bank;
      }
    }
    for (uint32_t i = 0; i < n; i++) seq[b + i] = order[i];
  }

  // ---- (2) local search over the bank groups ---------------------------------------------------------------------
  // A group's results keep distinct bank groups; two of them may trade theirs (or one may move to a bank group the
  // group does not use) whenever that lowers the conflicts of the groups that READ them. With every reader known this
  // removes most of what the level-by-level choice above leaves: the 1M-gate DAG goes from 1.10 to ~1.01 wavefronts
  // per ideal wavefront.
  if (!std::getenv("CLG_NO_BANK_SEARCH")) {
    // group id of every position; per (group, operand position) a histogram of bank groups read
    std::vector<uint32_t> gbase(n_levels + 1, 0);
    for (uint32_t l = 0; l < n_levels; l++) gbase[l + 1] = gbase[l] + (cnt[l + 1] - cnt[l] + BANKS - 1) / BANKS;
    const uint32_t G = gbase[n_levels];
    std::vector<uint8_t> hist((size_t)G * 3 * BANKS, 0);
    std::vector<uint32_t> roff(V + 1, 0), reads;  // per value: (group * 3 + position) of every read
    auto each_read = [&](auto &&fn) {
      for (uint32_t l = 0; l < n_levels; l++)
        for (uint32_t i = cnt[l]; i < cnt[l + 1]; i++) {
          const uint32_t x = seq[i].x, g = gbase[l] + (i - cnt[l]) / BANKS;
          if (x >= V) {
            if (net.outs[x - V].val != NONE) fn(net.outs[x - V].val, g * 3);
          } else if (net.vals[x].is_lut) {
            const uint8_t *pm = PERM[seq[i].perm];
            for (int k = 0; k < net.vals[x].nfan; k++) fn(net.vals[x].fan[k], g * 3 + pm[k]);
          }
        }
    };
    each_read([&](uint32_t v, uint32_t) { roff[v + 1]++; });
    for (size_t v = 0; v < V; v++) roff[v + 1] += roff[v];
    reads.resize(roff[V]);
    {
      std::vector<uint32_t> fill(roff.begin(), roff.end() - 1);
      each_read([&](uint32_t v, uint32_t where) { reads[fill[v]++] = where, hist[(size_t)where * BANKS + bankv[v]]++; });
    }
    // cost of moving value v from bank group `from` to `to`: change of (sum of squared counts) over its readers' groups
    auto delta = [&](uint32_t v, uint32_t from, uint32_t to) {
      int d = 0;
      for (uint32_t r = roff[v]; r < roff[v + 1]; r++) {
        const uint8_t *h = &hist[(size_t)reads[r] * BANKS];
        d += 2 * ((int)h[to] - (int)h[from]) + 2;
      }
      return d;
    };
    auto move = [&](uint32_t v, uint32_t from, uint32_t to) {
      for (uint32_t r = roff[v]; r < roff[v + 1]; r++) {
        uint8_t *h = &hist[(size_t)reads[r] * BANKS];
        h[from]--, h[to]++;
      }
      bankv[v] = (uint8_t)to;
    };
    for (int sweep = 0; sweep < 8; sweep++) {
      uint64_t improved = 0;
      for (uint32_t l = 0; l < n_levels; l++)
        for (uint32_t i0 = cnt[l]; i0 < cnt[l + 1]; i0 += BANKS) {
          uint32_t mem[MAX_BANKS], m = 0, used = 0;
          for (uint32_t i = i0; i < std::min(cnt[l + 1], i0 + BANKS); i++)
            if (seq[i].x < V) mem[m++] = seq[i].x, used |= 1u << bankv[seq[i].x];
          for (uint32_t a = 0; a < m; a++) {
            const uint32_t va = mem[a];
            if (roff[va] == roff[va + 1]) continue;
            // to a bank group nobody in the group uses
            for (uint32_t to = 0; to < BANKS; to++) {
              if (used >> to & 1u) continue;
              const uint32_t from = bankv[va];
              if (delta(va, from, to) < 0) move(va, from, to), used = (used & ~(1u << from)) | 1u << to, improved++;
            }
            // or trade with another member
            for (uint32_t c = a + 1; c < m; c++) {
              const uint32_t vc = mem[c], ba = bankv[va], bc = bankv[vc];
              // (the two moves interact only if both values are read by the same group and position: apply, measure, undo)
              const int d1 = delta(va, ba, bc);
              move(va, ba, bc);
              const int d2 = delta(vc, bc, ba);
              if (d1 + d2 < 0) move(vc, bc, ba), improved++;
              else move(va, bc, ba);
            }
          }
        }
      if (std::getenv("CLG_TIMING")) {
        uint64_t cost = 0, wf = 0, ideal = 0;
        for (size_t q = 0; q < (size_t)G * 3; q++) {
          uint32_t mx = 0;
          for (uint32_t k = 0; k < BANKS; k++) cost += (uint64_t)hist[q * BANKS + k] * hist[q * BANKS + k], mx = std::max<uint32_t>(mx, hist[q * BANKS + k]);
          wf += mx, ideal += mx > 0;
        }
        std::fprintf(stderr, "[clg]   bank search sweep %d: %llu moves, sum of squares %llu, wavefronts %llu / ideal %llu\n", sweep, (unsigned long long)improved,
                     (unsigned long long)cost, (unsigned long long)wf, (unsigned long long)ideal);
      }
      if (!improved) break;
    }
  }

  // ---- (3) slots: a result takes a free slot of its bank group; a slot freed by its last reader in level L is handed
  // out again from level L + 1 on
  for (uint32_t l = 0; l < n_levels; l++) {
    if (l > 0) {
      for (uint32_t sl : release[l - 1]) pool.put(sl);
      release[l - 1].clear(), release[l - 1].shrink_to_fit();
    }
    const uint32_t b = cnt[l], n = cnt[l + 1] - cnt[l];
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t x = seq[b + i].x;
      FlatInstr in{};
      if (x >= V) {
        const OutRef &o = net.outs[x - V];
        in.row = o.row, in.flags = o.inv ? 1 : 0;
        if (o.val == NONE) in.kind = K_STCONST;
        else in.kind = K_STOUT, in.a = (uint16_t)slot[o.val];
      } else {
        const Val &v = net.vals[x];
        if (v.is_lut) {
          in = make_lut(v);
          const uint8_t *pm = PERM[seq[b + i].perm];
          uint32_t sl[3] = {NONE, NONE, NONE};
          uint8_t where[3] = {0, 1, 2};
          for (int k = 0; k < v.nfan; k++) sl[pm[k]] = slot[v.fan[k]], where[k] = pm[k];
          for (int k = 0; k < 3; k++)
            if (sl[k] == NONE) sl[k] = sl[where[0]];  // positions the LUT does not read alias a read one
          // (the permutation is one of the function's symmetries, so the table is unchanged)
          in.a = (uint16_t)sl[0], in.b = (uint16_t)sl[1], in.c = (uint16_t)sl[2];
        } else {
          in.kind = K_LDIN, in.row = v.row;
        }
        slot[x] = pool.get(bankv[x]);
        if (slot[x] > 0xFFFF) {
          s.ok = false;
          return s;
        }
        in.dst = (uint16_t)slot[x];
        release[last[x]].push_back(slot[x]);
      }
      s.instr[b + i] = in;
    }
  }
  s.n_slots = pool.high;
  return s;
}

// ---- row stream (clg_coop_rows; format in flat.hpp, FlatRows) ---------------------------------------------------
// A forward list schedule over segments. Segment s may read from shared memory what was produced in segments
// <= s - LAG; at most `cap_rows` rows (of 32 LUTs, all with the same operand count and at most three truth tables) are
// issued per segment, most urgent LUTs first, dealt to the warps round-robin. A lane that still holds the result of
// its previous LUT in registers prefers a LUT that reads it (operand a then comes from the register: one shared-memory
// load less; a value that is only ever read this way is never stored). Free lanes are filled from the ready LUTs of
// the row's tables so that the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups per operand.
struct RowsSched {
  bool ok = false;
  uint32_t n_slots = 0, n_warps = 0, n_segs = 0, lag = 2, n_from_reg = 0, n_no_store = 0;
  std::vector<uint32_t> rowtab, sidetab;
  std::vector<FlatInstr> instr, side;
};

RowsSched schedule_rows(const Network &net, uint32_t NW, uint32_t cap_rows) {
  constexpr uint32_t LAG = 2, LANES = 32, BANKS = 8;
  const size_t V = net.vals.size(), O = net.outs.size();
  RowsSched rs;
  rs.n_warps = NW, rs.lag = LAG;
  size_t n_luts = 0, n_loads = 0;
  for (size_t v = 0; v < V; v++) n_luts += net.vals[v].is_lut, n_loads += !net.vals[v].is_lut;
  if (n_luts == 0 || n_loads > 4096) return rs;  // (every input is loaded in segment 0 and stays live until its last reader)

  // priority: unit-delay ALAP level (small = urgent)
  std::vector<uint32_t> asap(V, 0), alap(V, NONE);
  uint32_t depth = 0;
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut) {
      uint32_t h = 0;
      for (int i = 0; i < net.vals[v].nfan; i++) h = std::max(h, asap[net.vals[v].fan[i]]);
      asap[v] = h + 1, depth = std::max(depth, h + 1);
    }
  for (const OutRef &o : net.outs)
    if (o.val != NONE) alap[o.val] = depth;
  for (size_t v = V; v-- > 0;) {
    if (alap[v] == NONE) alap[v] = net.vals[v].is_lut ? asap[v] : 0;
    if (net.vals[v].is_lut)
      for (int i = 0; i < net.vals[v].nfan; i++) {
        uint32_t &a = alap[net.vals[v].fan[i]];
        a = std::min(a, alap[v] - 1);
      }
  }
  // readers of each value
  std::vector<uint32_t> coff(V + 1, 0), cons;
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut)
      for (int i = 0; i < net.vals[v].nfan; i++) coff[net.vals[v].fan[i] + 1]++;
  for (size_t v = 0; v < V; v++) coff[v + 1] += coff[v];
  cons.resize(coff[V]);
  {
    std::vector<uint32_t> fill(coff.begin(), coff.end() - 1);
    for (size_t v = 0; v < V; v++)
      if (net.vals[v].is_lut)
        for (int i = 0; i < net.vals[v].nfan; i++) cons[fill[net.vals[v].fan[i]]++] = (uint32_t)v;
  }
  static const uint8_t PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};

  // what a lane executes
  struct LaneOp {
    uint32_t v = NONE;     // value (LUT) or NONE: no instruction
    uint8_t pos[3] = {0, 1, 2};  // operand k of the value sits at position pos[k]
    uint8_t tt = 0;
    bool from_reg = false;  // the operand at position 0 is the lane's previous result
  };
  struct Row {
    LaneOp lane[LANES];
  };
  std::vector<std::vector<Row>> rows;   // per (segment * NW + warp)
  std::vector<uint32_t> seg_of(V, NONE);
  std::vector<uint8_t> vis(V, 0), miss(V, 0), bankv(V, 0), reg_read(V, 0);
  std::vector<uint32_t> smem_reads(V, 0), last_read(V, 0);  // shared-memory reads of each value; last segment that reads it
  std::vector<std::vector<uint32_t>> produced;              // values produced per segment
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut) miss[v] = net.vals[v].nfan;
  using HeapItem = std::pair<uint32_t, uint32_t>;  // (alap, value)
  std::vector<std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>>> ready(256);
  std::vector<uint32_t> held(NW * LANES, NONE);
  uint64_t bank_weight[BANKS] = {};
  size_t left = n_luts;

  // inputs: loaded in segment 0 through the side list
  produced.emplace_back();
  for (size_t v = 0; v < V; v++)
    if (!net.vals[v].is_lut) {
      seg_of[v] = 0, produced[0].push_back((uint32_t)v);
      uint32_t b = 0;
      for (uint32_t k = 1; k < BANKS; k++)
        if (bank_weight[k] < bank_weight[b]) b = k;
      bankv[v] = (uint8_t)b, bank_weight[b] += coff[v + 1] - coff[v];
    }

  auto tt_of = [&](const Val &x, const uint8_t pos[3]) { return remap(x.tt, pos); };
  auto clean = [&](uint32_t t) {
    while (!ready[t].empty() && seg_of[ready[t].top().second] != NONE) ready[t].pop();
  };
  uint32_t start_w = 0;
  uint64_t dbg_picks = 0, dbg_conf = 0, dbg_win = 0;
  uint32_t min_reg_quarter = 7;
  if (const char *env = std::getenv("CLG_ROWS_MIN_REG")) min_reg_quarter = (uint32_t)std::atoi(env);  // tuning knob (1: every candidate, 9: none)
  struct WinItemT {
    uint32_t v;
    uint8_t nops, perms, bank[3];
    bool taken;
  };
  std::vector<WinItemT> window;
  struct SigRef {
    size_t c;
    int pq;
  };
  std::vector<std::vector<SigRef>> by_sig(512);
  for (uint32_t sgm = 0; left > 0; sgm++) {
    if (sgm > 4 * (uint32_t)V + 64) return rs;  // (cannot happen: every segment with nothing to do is followed by one with something)
    if (produced.size() <= sgm) produced.resize(sgm + 1);
    rows.resize((size_t)(sgm + 1) * NW);
    if (sgm >= LAG)
      for (uint32_t u : produced[sgm - LAG]) {
        vis[u] = 1;
        for (uint32_t k = coff[u]; k < coff[u + 1]; k++) {
          const uint32_t c = cons[k];
          if (seg_of[c] == NONE && --miss[c] == 0) ready[net.vals[c].tt].push({alap[c], c});
        }
      }
    uint32_t n_rows = 0, idle = 0, w = start_w;
    while (n_rows < cap_rows && idle < NW) {
      // ---- one row for warp w
      // per lane: the most urgent unscheduled reader of the held value whose OTHER operands are all visible
      uint32_t cand[LANES];
      LaneOp cform[LANES];
      uint32_t count[256] = {};
      for (uint32_t i = 0; i < LANES; i++) {
        cand[i] = NONE;
        const uint32_t h = held[w * LANES + i];
        if (h == NONE) continue;
        uint32_t best = NONE;
        for (uint32_t k = coff[h]; k < coff[h + 1]; k++) {
          const uint32_t c = cons[k];
          if (seg_of[c] != NONE) continue;
          if (!(miss[c] == 0 || (miss[c] == 1 && !vis[h]))) continue;
          if (best == NONE || alap[c] < alap[best]) best = c;
        }
        if (best == NONE) continue;
        // operand order: the held value first, the others in the order that gives the smaller table
        const Val &x = net.vals[best];
        int ci = 0;
        while (x.fan[ci] != h) ci++;
        LaneOp f;
        f.v = best, f.from_reg = true;
        uint8_t p0[3] = {0, 1, 2}, p1[3] = {0, 1, 2};
        int others[2], no = 0;
        for (int k = 0; k < x.nfan; k++)
          if (k != ci) others[no++] = k;
        p0[ci] = 0, p1[ci] = 0;
        if (no >= 1) p0[others[0]] = 1, p1[others[0]] = (uint8_t)(no == 2 ? 2 : 1);
        if (no == 2) p0[others[1]] = 2, p1[others[1]] = 1;
        for (int k = x.nfan; k < 3; k++) p0[k] = p1[k] = (uint8_t)k;
        const uint8_t t0 = tt_of(x, p0), t1 = tt_of(x, p1);
        if (t1 < t0) std::memcpy(f.pos, p1, 3), f.tt = t1;
        else std::memcpy(f.pos, p0, 3), f.tt = t0;
        // only forwardings that keep the LUT's own table (the held value is, or may trade places with, its operand a): a
        // second table in the row costs a predicated LOP3 pass, more than the saved load is worth
        if (f.tt != x.tt && !std::getenv("CLG_ROWS_ANY_TT")) continue;
        cand[i] = best, cform[i] = f;
        count[f.tt]++;
      }
      // the row's operand count and tables: register candidates first, then the most urgent ready LUTs
      uint32_t score[4] = {0, 0, 0, 0}, urgent[4] = {NONE, NONE, NONE, NONE};
      for (uint32_t t = 0; t < 256; t++) {
        clean(t);
        // (a table names its operand count: a function of two inputs does not depend on c)
      }
      auto nf_of_tt = [&](uint32_t t) {  // operands a table really depends on
        const uint8_t x = (uint8_t)t;
        const bool da = ((x >> 4) & 0x0F) != (x & 0x0F), db = ((x >> 2) & 0x33) != (x & 0x33), dc = ((x >> 1) & 0x55) != (x & 0x55);
        return (uint32_t)da + db + dc;
      };
      for (uint32_t i = 0; i < LANES; i++)
        if (cand[i] != NONE) score[net.vals[cand[i]].nfan] += 2;
      for (uint32_t t = 0; t < 256; t++)
        if (!ready[t].empty()) {
          const uint32_t nf = net.vals[ready[t].top().second].nfan;
          score[nf] += (uint32_t)std::min<size_t>(ready[t].size(), LANES);
          urgent[nf] = std::min(urgent[nf], ready[t].top().first);
        }
      (void)nf_of_tt;
      int cls = -1;
      for (int nf = 1; nf <= 3; nf++)
        if (score[nf] && (cls < 0 || urgent[nf] < urgent[cls] || (urgent[nf] == urgent[cls] && score[nf] > score[cls]))) cls = nf;
      // (the most urgent ready LUT decides, so that no operand count starves; register candidates alone also make a row)
      if (cls < 0) {
        idle++, w = (w + 1) % NW;
        continue;
      }
      uint8_t tts[3];
      uint32_t n_tt = 0;
      for (; n_tt < 3; n_tt++) {  // register candidates' tables, most frequent first
        uint32_t best = 256;
        for (uint32_t t = 0; t < 256; t++)
          if (count[t] && (best == 256 || count[t] > count[best])) {
            bool cls_ok = false;
            for (uint32_t i = 0; i < LANES && !cls_ok; i++) cls_ok = cand[i] != NONE && cform[i].tt == t && net.vals[cand[i]].nfan == cls;
            if (cls_ok) best = t;
          }
        if (best == 256) break;
        tts[n_tt] = (uint8_t)best, count[best] = 0;
      }
      while (n_tt < 3) {  // then the tables of the most urgent ready LUTs of this operand count
        uint32_t best = 256;
        for (uint32_t t = 0; t < 256; t++) {
          if (ready[t].empty() || net.vals[ready[t].top().second].nfan != cls) continue;
          bool have = false;
          for (uint32_t k = 0; k < n_tt; k++) have |= tts[k] == t;
          if (!have && (best == 256 || ready[t].top().first < ready[best].top().first)) best = t;
        }
        if (best == 256) break;
        tts[n_tt++] = (uint8_t)best;
      }
      auto in_set = [&](uint8_t t) {
        for (uint32_t k = 0; k < n_tt; k++)
          if (tts[k] == t) return true;
        return false;
      };
      // ---- fill the lanes, quarter by quarter
      // free lanes are filled from a WINDOW of the most urgent ready LUTs of the row's tables (taken off the heaps,
      // what is left goes back): a lane takes the first one whose operands, in one of the orders its table allows,
      // fall into bank groups the quarter does not use yet
      Row row;
      uint32_t filled = 0, from_reg = 0;
      using WinItem = WinItemT;
      std::vector<WinItem> &win = window;
      win.clear();
      {
        static const uint32_t WIN = std::getenv("CLG_ROWS_WIN") ? (uint32_t)std::atoi(std::getenv("CLG_ROWS_WIN")) : 1024;  // (knob)
        while (win.size() < WIN) {
          uint32_t bt = 256;
          for (uint32_t k = 0; k < n_tt; k++) {
            clean(tts[k]);
            if (!ready[tts[k]].empty() && net.vals[ready[tts[k]].top().second].nfan == cls &&
                (bt == 256 || ready[tts[k]].top().first < ready[bt].top().first))
              bt = tts[k];
          }
          if (bt == 256) break;
          const uint32_t v = ready[bt].top().second;
          ready[bt].pop();
          const Val &x = net.vals[v];
          WinItem it{v, x.nfan, 1, {0, 0, 0}, false};
          for (int k = 0; k < x.nfan; k++) it.bank[k] = bankv[x.fan[k]];
          if (x.nfan >= 2)
            for (int pq = 1; pq < 6; pq++) {
              bool fits = true;
              for (int k = 0; k < x.nfan; k++) fits &= PERM[pq][k] < x.nfan;
              if (fits && remap(x.tt, PERM[pq]) == x.tt) it.perms |= (uint8_t)(1u << pq);
            }
          win.push_back(it);
        }
      }
      size_t win_head = 0;
      for (auto &l : by_sig) l.clear();
      for (size_t c = 0; c < win.size(); c++)
        for (int pq = 0; pq < 6; pq++) {
          if (!(win[c].perms >> pq & 1)) continue;
          uint32_t sig = 0;
          for (int k = 0; k < win[c].nops; k++) sig |= (uint32_t)win[c].bank[k] << (3 * PERM[pq][k]);
          by_sig[sig].push_back({c, pq});
        }
      for (uint32_t q0 = 0; q0 < LANES; q0 += BANKS) {
        uint32_t mask[3] = {0, 0, 0};
        // An LDS.128 serves a quarter-warp at a time, so a register operand only saves a wavefront when (nearly) the
        // whole quarter has one: a quarter takes its register candidates if at least 7 of its 8 lanes have one (the
        // eighth stays empty), otherwise none of them (scattered ones save nothing and only pin LUTs to lanes where
        // their other operands collide).
        uint32_t n_cand = 0;
        for (uint32_t i = q0; i < q0 + BANKS; i++)
          n_cand += cand[i] != NONE && net.vals[cand[i]].nfan == (uint32_t)cls && in_set(cform[i].tt) && seg_of[cand[i]] == NONE;
        const bool reg_quarter = n_cand >= min_reg_quarter;
        // register candidates are where they are
        for (uint32_t i = q0; i < q0 + BANKS; i++) {
          if (!reg_quarter) break;
          if (cand[i] == NONE || net.vals[cand[i]].nfan != cls || !in_set(cform[i].tt) || seg_of[cand[i]] != NONE) continue;
          const Val &x = net.vals[cand[i]];
          LaneOp f = cform[i];
          // its two loaded operands may trade places if the table allows it and that avoids a conflict
          auto conflicts = [&](const uint8_t pos[3]) {
            uint32_t n = 0;
            for (int k = 0; k < x.nfan; k++)
              if (pos[k] != 0) n += mask[pos[k]] >> bankv[x.fan[k]] & 1u;
            return n;
          };
          if (x.nfan == 3 && conflicts(f.pos)) {
            uint8_t alt[3];
            for (int k = 0; k < 3; k++) alt[k] = f.pos[k] == 0 ? 0 : (uint8_t)(3 - f.pos[k]);
            if (tt_of(x, alt) == f.tt && conflicts(alt) == 0) std::memcpy(f.pos, alt, 3);
          }
          for (int k = 0; k < x.nfan; k++)
            if (f.pos[k] != 0) mask[f.pos[k]] |= 1u << bankv[x.fan[k]];
          row.lane[i] = f, seg_of[cand[i]] = sgm, filled++, from_reg++;
        }
        if (reg_quarter) continue;
        // ---- a plain quarter: 8 LUTs from the window whose operands, position by position, hit 8 different bank groups.
        // The first six are taken first-fit in order of urgency (easy: most bank groups are still free); the last two
        // must have exactly the bank groups that are left, so they are looked up by signature; if no such pair is in the
        // window the sixth choice is varied. What still collides after that is placed with the fewest conflicts.
        {
          struct Pick {
            size_t c;
            int pq;
          };
          Pick picks[BANKS];
          uint32_t np = 0;
          auto fits = [&](const WinItem &it, int pq) {
            for (int k = 0; k < it.nops; k++)
              if (mask[PERM[pq][k]] >> it.bank[k] & 1u) return false;
            return true;
          };
          auto apply = [&](const WinItem &it, int pq, bool on) {
            for (int k = 0; k < it.nops; k++) {
              if (on) mask[PERM[pq][k]] |= 1u << it.bank[k];
              else mask[PERM[pq][k]] &= ~(1u << it.bank[k]);
            }
          };
          auto usable = [&](size_t c) { return !win[c].taken && seg_of[win[c].v] == NONE; };
          auto first_fit = [&](size_t from, size_t &found, int &found_pq) {
            for (size_t c = from; c < win.size(); c++) {
              if (!usable(c)) continue;
              for (int pq = 0; pq < 6; pq++)
                if ((win[c].perms >> pq & 1) && fits(win[c], pq)) {
                  found = c, found_pq = pq;
                  return true;
                }
            }
            return false;
          };
          while (win_head < win.size() && !usable(win_head)) win_head++;
          const uint32_t nops = (uint32_t)cls;
          // greedy part
          while (np < BANKS - 2) {
            size_t c;
            int pq;
            if (!first_fit(win_head, c, pq)) break;
            win[c].taken = true, apply(win[c], pq, true), picks[np++] = {c, pq};
          }
          // the last two by signature (key: bank group at each position)
          auto complete_pair = [&]() -> bool {
            uint32_t miss_b[3][2], nm[3] = {0, 0, 0};
            for (uint32_t pos = 0; pos < nops; pos++)
              for (uint32_t bk = 0; bk < BANKS; bk++)
                if (!(mask[pos] >> bk & 1u)) {
                  if (nm[pos] == 2) return false;
                  miss_b[pos][nm[pos]++] = bk;
                }
            for (uint32_t pos = 0; pos < nops; pos++)
              if (nm[pos] != 2) return false;
            for (uint32_t combo = 0; combo < (1u << nops) / 2 + (nops == 0); combo++) {  // x takes miss[pos][bit], y the other; x <-> y symmetric
              uint32_t sx = 0, sy = 0;
              for (uint32_t pos = 0; pos < nops; pos++) {
                const uint32_t bit = pos == 0 ? 0 : (combo >> (pos - 1)) & 1u;
                sx |= miss_b[pos][bit] << (3 * pos), sy |= miss_b[pos][bit ^ 1u] << (3 * pos);
              }
              size_t cx = win.size(), cy = win.size();
              int px = 0, py = 0;
              for (const SigRef &r : by_sig[sx])
                if (usable(r.c)) {
                  cx = r.c, px = r.pq;
                  break;
                }
              if (cx == win.size()) continue;
              for (const SigRef &r : by_sig[sy])
                if (usable(r.c) && r.c != cx) {
                  cy = r.c, py = r.pq;
                  break;
                }
              if (cy == win.size()) continue;
              win[cx].taken = win[cy].taken = true;
              apply(win[cx], px, true), apply(win[cy], py, true);
              picks[np++] = {cx, px}, picks[np++] = {cy, py};
              return true;
            }
            return false;
          };
          if (np == BANKS - 2 && !complete_pair()) {
            // vary the sixth choice, then the fifth and the sixth
            const Pick fifth = picks[np - 2], sixth = picks[np - 1];
            apply(win[sixth.c], sixth.pq, false), win[sixth.c].taken = false, np--;
            bool done = false;
            auto vary_last = [&](int attempts) {
              size_t from = win_head;
              for (int attempt = 0; attempt < attempts && !done; attempt++) {
                size_t c;
                int pq;
                if (!first_fit(from, c, pq)) break;
                from = c + 1;
                win[c].taken = true, apply(win[c], pq, true), picks[np++] = {c, pq};
                done = complete_pair();
                if (!done) apply(win[c], pq, false), win[c].taken = false, np--;
              }
            };
            vary_last(64);
            if (!done) {
              apply(win[fifth.c], fifth.pq, false), win[fifth.c].taken = false, np--;
              size_t from5 = win_head;
              for (int a5 = 0; a5 < 24 && !done; a5++) {
                size_t c;
                int pq;
                if (!first_fit(from5, c, pq)) break;
                from5 = c + 1;
                if (c == fifth.c) continue;
                win[c].taken = true, apply(win[c], pq, true), picks[np++] = {c, pq};
                vary_last(24);
                if (!done) apply(win[c], pq, false), win[c].taken = false, np--;
              }
              if (!done) win[fifth.c].taken = true, apply(win[fifth.c], fifth.pq, true), picks[np++] = fifth;
            }
            if (!done) {  // settle for the first choices and the fewest conflicts
              win[sixth.c].taken = true, apply(win[sixth.c], sixth.pq, true), picks[np++] = sixth;
              while (np < BANKS) {
                size_t best = win.size();
                int best_pq = 0;
                uint32_t best_n = NONE, seen = 0;
                for (size_t c = win_head; c < win.size() && seen < 256; c++) {
                  if (!usable(c)) continue;
                  seen++;
                  for (int pq = 0; pq < 6; pq++) {
                    if (!(win[c].perms >> pq & 1)) continue;
                    uint32_t n = 0;
                    for (int k = 0; k < win[c].nops; k++) n += mask[PERM[pq][k]] >> win[c].bank[k] & 1u;
                    if (n < best_n) best_n = n, best = c, best_pq = pq;
                  }
                }
                if (best == win.size()) break;
                dbg_conf++;
                win[best].taken = true, apply(win[best], best_pq, true), picks[np++] = {best, best_pq};
              }
            }
          }
          dbg_picks += np;
          for (uint32_t j = 0; j < np; j++) {
            const WinItem &it = win[picks[j].c];
            LaneOp f;
            f.v = it.v, f.tt = net.vals[it.v].tt, f.from_reg = false;
            for (int k = 0; k < 3; k++) f.pos[k] = k < it.nops ? PERM[picks[j].pq][k] : (uint8_t)k;
            row.lane[q0 + j] = f, seg_of[it.v] = sgm, filled++;
          }
        }
      }
      for (const WinItem &it : win)
        if (!it.taken && seg_of[it.v] == NONE) ready[net.vals[it.v].tt].push({alap[it.v], it.v});
      // a thin row early in the program is not worth its instruction fetch: leave its LUTs for later, unless it is
      // mostly register forwarding or little else is left
      if (filled == 0 || (filled < 8 && from_reg * 2 < filled && left > 4096 && n_rows > 0)) {
        for (uint32_t i = 0; i < LANES; i++)
          if (row.lane[i].v != NONE) {
            const uint32_t v = row.lane[i].v;
            seg_of[v] = NONE;
            if (miss[v] == 0) ready[net.vals[v].tt].push({alap[v], v});
          }
        idle++, w = (w + 1) % NW;
        continue;
      }
      idle = 0;
      // bookkeeping: reads, results, bank groups of the results (distinct within a quarter, least read traffic first)
      for (uint32_t q0 = 0; q0 < LANES; q0 += BANKS) {
        uint32_t idx[BANKS], m = 0, dmask = 0;
        for (uint32_t i = q0; i < q0 + BANKS; i++)
          if (row.lane[i].v != NONE) idx[m++] = i;
        std::stable_sort(idx, idx + m, [&](uint32_t a, uint32_t b) {
          return coff[row.lane[a].v + 1] - coff[row.lane[a].v] > coff[row.lane[b].v + 1] - coff[row.lane[b].v];
        });
        for (uint32_t j = 0; j < m; j++) {
          uint32_t bank = NONE;
          for (uint32_t c = 0; c < BANKS; c++)
            if (!(dmask >> c & 1u) && (bank == NONE || bank_weight[c] < bank_weight[bank])) bank = c;
          dmask |= 1u << bank;
          const uint32_t v = row.lane[idx[j]].v;
          bankv[v] = (uint8_t)bank, bank_weight[bank] += coff[v + 1] - coff[v];
        }
      }
      for (uint32_t i = 0; i < LANES; i++) {
        const LaneOp &f = row.lane[i];
        if (f.v == NONE) {
          held[w * LANES + i] = NONE;  // the row's LOP3s run in every lane: a lane without an instruction loses its register
          continue;
        }
        const Val &x = net.vals[f.v];
        for (int k = 0; k < x.nfan; k++) {
          if (f.from_reg && f.pos[k] == 0) reg_read[x.fan[k]] = 1;
          else smem_reads[x.fan[k]]++, last_read[x.fan[k]] = sgm;
        }
        held[w * LANES + i] = f.v;
        produced[sgm].push_back(f.v);
        left--;
        rs.n_from_reg += f.from_reg;
      }
      rows[(size_t)sgm * NW + w].push_back(row);
      n_rows++, w = (w + 1) % NW;
    }
    start_w = w;
  }
  uint32_t n_segs = (uint32_t)(rows.size() / NW);
  if (std::getenv("CLG_TIMING"))
    std::fprintf(stderr, "[clg]   rows: %llu window picks, %llu with a conflict, mean window %.0f\n", (unsigned long long)dbg_picks, (unsigned long long)dbg_conf,
                 dbg_picks ? (double)dbg_win / dbg_picks : 0.0);

  // ---- stores: LAG segments after the value exists (constants too: a constant may overwrite a carried row that is
  // loaded in segment 0)
  std::vector<std::vector<uint32_t>> side_of(n_segs + LAG + 1);  // per segment: V + out index, or value id (LDIN)
  for (size_t v = 0; v < V; v++)
    if (!net.vals[v].is_lut) side_of[0].push_back((uint32_t)v);
  for (size_t o = 0; o < O; o++) {
    const OutRef &r = net.outs[o];
    const uint32_t at = (r.val == NONE ? 0 : seg_of[r.val]) + LAG;
    side_of[at].push_back((uint32_t)(V + o));
    if (r.val != NONE) smem_reads[r.val]++, last_read[r.val] = std::max(last_read[r.val], at);
  }
  while (side_of.size() > 1 && side_of.back().empty()) side_of.pop_back();
  n_segs = std::max<uint32_t>(n_segs, (uint32_t)side_of.size());
  rows.resize((size_t)n_segs * NW);
  side_of.resize(n_segs);

  // ---- slots: a stored result takes a free slot of its bank group; a slot whose last reader runs in segment r may be
  // written again from segment r + LAG on
  std::vector<uint32_t> slot(V, NONE);
  std::vector<std::vector<uint32_t>> release(n_segs + LAG + 1);
  std::vector<uint32_t> free_[BANKS];
  uint32_t fresh[BANKS], high = 0;
  for (uint32_t b = 0; b < BANKS; b++) fresh[b] = b;
  auto take = [&](uint32_t v, uint32_t sgm) {
    const uint32_t b = bankv[v];
    uint32_t sl;
    if (!free_[b].empty()) sl = free_[b].back(), free_[b].pop_back();
    else sl = fresh[b], fresh[b] += BANKS;
    high = std::max(high, sl + 1);
    slot[v] = sl;
    release[std::max(last_read[v], sgm) + LAG].push_back(sl);
  };
  rs.rowtab.assign((size_t)n_segs * NW + 1, 0);
  rs.sidetab.assign(n_segs + 1, 0);
  for (uint32_t sgm = 0; sgm < n_segs; sgm++) {
    for (uint32_t sl : release[sgm]) free_[sl % BANKS].push_back(sl);
    for (uint32_t x : side_of[sgm])
      if (x < V) take(x, sgm);
    for (uint32_t w = 0; w < NW; w++)
      for (const Row &row : rows[(size_t)sgm * NW + w])
        for (uint32_t i = 0; i < LANES; i++)
          if (row.lane[i].v != NONE && smem_reads[row.lane[i].v]) take(row.lane[i].v, sgm);
    if (high > 16383) return rs;  // (the row kernel's instruction words hold 14-bit slot indices)
    // emit
    rs.sidetab[sgm] = (uint32_t)rs.side.size();
    for (uint32_t x : side_of[sgm]) {
      FlatInstr in{};
      if (x >= V) {
        const OutRef &o = net.outs[x - V];
        in.row = o.row, in.flags = o.inv ? 1 : 0;
        if (o.val == NONE) in.kind = K_STCONST;
        else in.kind = K_STOUT, in.a = (uint16_t)slot[o.val];
      } else {
        in.kind = K_LDIN, in.row = net.vals[x].row, in.dst = (uint16_t)slot[x];
      }
      rs.side.push_back(in);
    }
    for (uint32_t w = 0; w < NW; w++) {
      rs.rowtab[(size_t)sgm * NW + w] = (uint32_t)rs.instr.size();
      for (const Row &row : rows[(size_t)sgm * NW + w])
        for (uint32_t i = 0; i < LANES; i++) {
          const LaneOp &f = row.lane[i];
          FlatInstr in{};
          if (f.v == NONE) {
            in.kind = K_NOP;
          } else {
            const Val &x = net.vals[f.v];
            in = make_lut(x);
            in.tt = f.tt;
            uint32_t sl[3] = {NONE, NONE, NONE};
            for (int k = 0; k < x.nfan; k++)
              if (!(f.from_reg && f.pos[k] == 0)) sl[f.pos[k]] = slot[x.fan[k]];
            uint32_t any = NONE;
            for (int k = 0; k < 3; k++)
              if (sl[k] != NONE && any == NONE) any = sl[k];
            if (any == NONE) any = 0;  // (a LUT of one operand that comes from the register reads nothing)
            for (int k = 0; k < 3; k++)
              if (sl[k] == NONE) sl[k] = any;  // positions that are not loaded alias a loaded one
            in.a = (uint16_t)sl[0], in.b = (uint16_t)sl[1], in.c = (uint16_t)sl[2];
            if (f.from_reg) in.flags |= FLAG_FROM_REG;
            if (smem_reads[f.v]) in.dst = (uint16_t)slot[f.v];
            else in.flags |= FLAG_NO_STORE, in.dst = 0, rs.n_no_store++;
          }
          rs.instr.push_back(in);
        }
    }
  }
  rs.rowtab[(size_t)n_segs * NW] = (uint32_t)rs.instr.size();
  rs.sidetab[n_segs] = (uint32_t)rs.side.size();
  rs.n_slots = std::max<uint32_t>(high, 1), rs.n_segs = n_segs;
  rs.ok = true;
  (void)reg_read;
  return rs;
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Connected components of the value network (an output belongs to the component of its value), bin-packed
// largest-first into at most `max_parts` parts of similar LUT count. Returns part id per value / per output.
uint32_t partition(const Network &net, uint32_t max_parts, std::vector<uint32_t> &part_of_val, std::vector<uint32_t> &part_of_out) {
  const size_t V = net.vals.size();
  std::vector<uint32_t> parent(V);
  for (size_t v = 0; v < V; v++) parent[v] = (uint32_t)v;
  auto find = [&](uint32_t x) {
    while (parent[x] != x) x = parent[x] = parent[parent[x]];
    return x;
  };
  for (size_t v = 0; v < V; v++)
    if (net.vals[v].is_lut)
      for (int i = 0; i < net.vals[v].nfan; i++) {
        uint32_t a = find((uint32_t)v), b = find(net.vals[v].fan[i]);
        if (a != b) parent[a] = b;
      }
  // carried state: the load and the store of one state row must stay in one part (in-place update ordering)
  std::vector<uint32_t> state_in(net.n_state, NONE);
  for (size_t v = 0; v < V; v++)
    if (!net.vals[v].is_lut && net.vals[v].row >= net.n_inputs) state_in[net.vals[v].row - net.n_inputs] = (uint32_t)v;
  for (const OutRef &o : net.outs)
    if (o.row >= net.n_outputs && o.val != NONE && state_in[o.row - net.n_outputs] != NONE) {
      uint32_t a = find(o.val), b = find(state_in[o.row - net.n_outputs]);
      if (a != b) parent[a] = b;
    }
  std::vector<uint32_t> weight(V, 0), comp_ids;
  for (size_t v = 0; v < V; v++) weight[find((uint32_t)v)] += 1;
  for (size_t v = 0; v < V; v++)
    if (find((uint32_t)v) == v) comp_ids.push_back((uint32_t)v);
  std::sort(comp_ids.begin(), comp_ids.end(), [&](uint32_t a, uint32_t b) { return weight[a] != weight[b] ? weight[a] > weight[b] : a < b; });
  const uint32_t P = (uint32_t)std::min<size_t>(max_parts, std::max<size_t>(1, comp_ids.size()));
  std::priority_queue<std::pair<uint64_t, uint32_t>, std::vector<std::pair<uint64_t, uint32_t>>, std::greater<std::pair<uint64_t, uint32_t>>> bins;
  for (uint32_t b = 0; b < P; b++) bins.push({0, b});
  std::vector<uint32_t> part_of_comp(V, 0);
  for (uint32_t c : comp_ids) {
    auto [load, b] = bins.top();
    bins.pop();
    part_of_comp[c] = b;
    bins.push({load + weight[c], b});
  }
  part_of_val.resize(V);
  for (size_t v = 0; v < V; v++) part_of_val[v] = part_of_comp[find((uint32_t)v)];
  part_of_out.resize(net.outs.size());
  for (size_t o = 0; o < net.outs.size(); o++) {
    const OutRef &r = net.outs[o];
    if (r.val != NONE) part_of_out[o] = part_of_val[r.val];
    else if (r.row >= net.n_outputs && state_in[r.row - net.n_outputs] != NONE)
      part_of_out[o] = part_of_val[state_in[r.row - net.n_outputs]];  // a constant stored over a carried row waits for that row's load
    else part_of_out[o] = 0;
  }
  return P;
}

Network subnetwork(const Network &net, const std::vector<uint32_t> &part_of_val, const std::vector<uint32_t> &part_of_out, uint32_t part) {
  Network sub;
  sub.n_inputs = net.n_inputs, sub.n_state = net.n_state, sub.n_outputs = net.n_outputs;
  std::vector<uint32_t> local(net.vals.size(), NONE);
  for (size_t v = 0; v < net.vals.size(); v++)
    if (part_of_val[v] == part) {
      Val x = net.vals[v];
      if (x.is_lut)
        for (int i = 0; i < x.nfan; i++) x.fan[i] = local[x.fan[i]];
      local[v] = (uint32_t)sub.vals.size();
      sub.vals.push_back(x);
    }
  for (size_t o = 0; o < net.outs.size(); o++)
    if (part_of_out[o] == part) {
      OutRef r = net.outs[o];
      if (r.val != NONE) r.val = local[r.val];
      sub.outs.push_back(r);
    }
  return sub;
}

}  // namespace

Flat levelize(const Simulation &sim, size_t n_immediates, const uint64_t *out_regs, size_t n_out, uint32_t flags) {
  const size_t R = sim.registers.size();
  const bool stateful = flags & CLG_LV_STATEFUL;
  if (R >= 0x7FFFFFFFu || sim.ops.size() >= 0x7FFFFFFFu) fail(CLG_E_UNSUPPORTED, "program too large");
  uint64_t n_nand = 0, n_set = 0;
  for (const Op &op : sim.ops) {
    if (op.tag == CLG_OP_NAND) {
      n_nand++;
      if (op.a >= R || op.b >= R || op.c >= R)
        fail(CLG_E_BAD_REGISTER, "Nand register out of range (the reference panics at simulation.rs:20-23)");
    } else if (op.tag == CLG_OP_SET) {
      n_set++;
      if (op.a >= R) fail(CLG_E_BAD_REGISTER, "Set register out of range (the reference panics at simulation.rs:26)");
    } else
      fail(CLG_E_INVALID, "unknown op tag");
  }
  std::vector<uint64_t> outs;
  if (out_regs == nullptr && n_out == 0) {
    outs.resize(R);
    for (size_t r = 0; r < R; r++) outs[r] = r;
  } else {
    outs.assign(out_regs, out_regs + n_out);
    for (uint64_t r : outs)
      if (r >= R) fail(CLG_E_BAD_REGISTER, "output register out of range (simulation.rs:34)");
  }
  if (n_immediates > 0x7FFFFFFFu) fail(CLG_E_INVALID, "too many immediates");

  // carried registers: first access in program order is a read (or an output nobody writes)
  std::vector<uint64_t> state_regs;
  if (stateful) {
    std::vector<uint8_t> first(R, 0);  // 1 = read first, 2 = written first
    for (const Op &op : sim.ops) {
      if (op.tag == CLG_OP_NAND) {
        if (!first[op.a]) first[op.a] = 1;
        if (!first[op.b]) first[op.b] = 1;
        if (!first[op.c]) first[op.c] = 2;
      } else if (!first[op.a])
        first[op.a] = 2;
    }
    for (uint64_t r : outs)
      if (!first[r]) first[r] = 1;
    for (size_t r = 0; r < R; r++)
      if (first[r] == 1) state_regs.push_back(r);
  }

  Phase ph_total("levelize total");
  std::unique_ptr<Phase> ph(new Phase("ssa + aig"));
  Aig g;
  std::vector<uint32_t> imm_lit(n_immediates);
  for (size_t i = 0; i < n_immediates; i++) imm_lit[i] = g.input((uint32_t)i);
  std::vector<uint32_t> cur(R, 0);
  for (size_t k = 0; k < state_regs.size(); k++) cur[state_regs[k]] = g.input((uint32_t)(n_immediates + k));
  for (const Op &op : sim.ops) {
    if (op.tag == CLG_OP_NAND) cur[op.c] = g.and_(cur[op.a], cur[op.b]) ^ 1u;
    else cur[op.a] = (op.a < n_immediates) ? imm_lit[op.a] : (op.b ? 1u : 0u);
  }
  std::vector<uint32_t> out_lits;
  for (uint64_t r : outs) out_lits.push_back(cur[r]);
  for (uint64_t r : state_regs) out_lits.push_back(cur[r]);

  ph.reset();
  Mapper m(g, out_lits, !(flags & CLG_LV_NO_FUSE));
  m.run();
  ph.reset(new Phase("network"));

  // build the value network: used input rows first, then LUT roots in topological order
  Network net;
  net.n_inputs = (uint32_t)n_immediates, net.n_state = (uint32_t)state_regs.size(), net.n_outputs = (uint32_t)outs.size();
  std::vector<uint32_t> cref = m.cover_refs();
  std::vector<uint32_t> val_of(g.size(), NONE);
  for (uint32_t n = 1; n < g.size(); n++)
    if (cref[n] && g.is_input(n)) {
      Val v{};
      v.row = g.f1[n];
      val_of[n] = (uint32_t)net.vals.size();
      net.vals.push_back(v);
    }
  uint32_t n_luts = 0;
  for (uint32_t n = 1; n < g.size(); n++)
    if (cref[n] && g.is_and(n)) {
      const Cut &c = m.best[n];
      Val v{};
      v.is_lut = 1, v.tt = c.tt, v.nfan = c.n;
      for (int i = 0; i < c.n; i++) v.fan[i] = val_of[c.leaf[i]];
      canonical_operand_order(v);
      val_of[n] = (uint32_t)net.vals.size();
      net.vals.push_back(v);
      n_luts++;
    }
  for (size_t o = 0; o < out_lits.size(); o++) {
    uint32_t l = out_lits[o];
    OutRef r{(uint32_t)o, NONE, (bool)(l & 1)};
    if ((l >> 1) != 0) r.val = val_of[l >> 1];
    // a carried register whose value is unchanged needs no store at all
    if (o >= outs.size() && r.val != NONE && !r.inv && !net.vals[r.val].is_lut &&
        net.vals[r.val].row == n_immediates + (o - outs.size()))
      continue;
    net.outs.push_back(r);
  }

  if (const char *path = std::getenv("CLG_DUMP_NETWORK")) {  // analysis knob: the mapped network as raw u32 records
    if (FILE *fp = std::fopen(path, "wb")) {
      const uint32_t hdr[4] = {(uint32_t)net.vals.size(), (uint32_t)net.outs.size(), net.n_inputs, net.n_outputs};
      std::fwrite(hdr, 4, 4, fp);
      for (const Val &v : net.vals) {
        const uint32_t rec[6] = {v.is_lut, v.tt, v.nfan, v.nfan > 0 ? v.fan[0] : v.row, v.fan[1], v.fan[2]};
        std::fwrite(rec, 4, 6, fp);
      }
      for (const OutRef &o : net.outs) {
        const uint32_t rec[3] = {o.row, o.val, o.inv};
        std::fwrite(rec, 4, 3, fp);
      }
      std::fclose(fp);
    }
  }
  ph.reset(new Phase("schedule: level"));
  // pack for the widest register-stack word count K that will fit one SM's shared memory (227 KB): K = 4 -> 8
  // bank groups of 16 B, K = 2 -> 16 of 8 B, K = 1 -> 32 banks of 4 B
  auto schedule_for_smem = [](const Network &n) {
    if (const char *env = std::getenv("CLG_LEVEL_BANKS")) return schedule_level(n, (uint32_t)std::atoi(env));  // tuning knob
    // the bank count follows from the K the stack will fit with, i.e. from the slot count, which is only known after
    // scheduling: start from the peak live count (slots end up 2-4 % above it) instead of trying 8, 16 and 32 in turn
    const size_t cap = 227 * 1024;
    auto banks_for = [&](double slots) { return slots * 16 <= cap ? 8u : slots * 8 <= cap ? 16u : 32u; };
    const uint32_t guess = banks_for(schedule_level(n, 8, true).n_slots * 1.06);
    Sched sc = schedule_level(n, guess);
    if (sc.ok && banks_for(sc.n_slots) != guess) sc = schedule_level(n, banks_for(sc.n_slots));  // off near a boundary
    return sc;
  };
  Sched level = schedule_for_smem(net);
  ph.reset(new Phase("schedule: lane"));
  Sched lane;
  lane.ok = false;
  if (!(flags & CLG_LV_NO_LANE_STREAM)) lane = schedule_lane(net);
  // a very wide, shallow program (10^5 latches side by side) can exceed the 65536 slots of the level order and
  // still need only a handful in lane order: the buffer then carries the lane stream alone
  if (!level.ok && !lane.ok) fail(CLG_E_UNSUPPORTED, "program needs more than 65536 physical registers in level order and in lane order");

  // partitioned level streams for small batches of wide, separable programs (e.g. flattened instances)
  ph.reset(new Phase("schedule: parts"));
  std::vector<Sched> parts;
  if (!(flags & CLG_LV_NO_PARTS) && n_luts >= 4096 && level.ok) {
    std::vector<uint32_t> pv, po;
    const uint32_t P = partition(net, 592, pv, po);
    if (P > 1)
      for (uint32_t q = 0; q < P; q++) {
        parts.push_back(schedule_for_smem(subnetwork(net, pv, po, q)));
        if (!parts.back().ok) {
          parts.clear();
          break;
        }
      }
  }

  // row stream for the cooperative row kernel: programs whose live set is beyond the specialised mode's reach and whose
  // level-ordered stack fits shared memory at four words per slot
  ph.reset(new Phase("schedule: rows"));
  RowsSched rowsch;
  const bool rows_forced = std::getenv("CLG_ROWS_FORCE") != nullptr;  // (tests: a row stream for any program)
  if (!(flags & CLG_LV_NO_ROWS) && level.ok && (size_t)level.n_slots * 16 <= 227 * 1024 &&
      (rows_forced || (n_luts >= 2048 && (!lane.ok || lane.n_slots > 400)))) {
    uint32_t depth = level.levels.empty() ? 1 : (uint32_t)level.levels.size() - 1;
    // warps of the CTA: enough that a segment of average width gives each of them a row or two
    uint32_t nw = (uint32_t)std::max<uint64_t>(2, std::min<uint64_t>(32, ((uint64_t)n_luts / depth + 47) / 48));
    if (const char *env = std::getenv("CLG_ROWS_WARPS")) nw = (uint32_t)std::max(1, std::min(32, std::atoi(env)));
    // rows per segment: about as many as the program's parallelism sustains when every non-forwarded dependence costs
    // two segments (more starves the scheduler of ready LUTs and leaves lanes empty, fewer only adds segments)
    uint32_t cap = (uint32_t)std::max<uint64_t>(nw / 2, std::min<uint64_t>(4 * nw, (uint64_t)n_luts / 32 * 100 / (195ull * depth)));  // (dag1m: 44; 41..45 measured equal, 47+ slower)
    if (const char *env = std::getenv("CLG_ROWS_CAP")) cap = (uint32_t)std::max(1, std::atoi(env));
    rowsch = schedule_rows(net, nw, cap);
    if (std::getenv("CLG_TIMING") && rowsch.ok)
      std::fprintf(stderr, "[clg]   rows: %u warps, cap %u rows/segment -> %u segments, %zu rows (%.1f %% of lanes empty), %u slots, %.1f %% from register, %.1f %% never stored\n",
                   nw, cap, rowsch.n_segs, rowsch.instr.size() / 32, 100.0 - 100.0 * n_luts / std::max<size_t>(1, rowsch.instr.size()), rowsch.n_slots,
                   100.0 * rowsch.n_from_reg / n_luts, 100.0 * rowsch.n_no_store / n_luts);
  }

  // serialise
  ph.reset(new Phase("serialise"));
  Flat f;
  size_t off = align16(sizeof(FlatHeader));
  FlatHeader h{};
  h.magic = FLAT_MAGIC, h.version = FLAT_VERSION, h.header_bytes = (uint32_t)sizeof(FlatHeader);
  h.n_inputs = net.n_inputs, h.n_outputs = net.n_outputs, h.n_state = net.n_state, h.n_luts = n_luts;
  h.ref_nand_count = n_nand, h.ref_set_count = n_set, h.flags = flags, h.n_registers = (uint32_t)R;
  h.off_state_regs = (uint32_t)off, off = align16(off + state_regs.size() * 8);
  h.off_out_regs = (uint32_t)off, off = align16(off + outs.size() * 8);
  const Sched *ss[2] = {&level, &lane};
  for (int i = 0; i < 2; i++) {
    if (!ss[i]->ok) continue;
    FlatStream &st = h.stream[i];
    st.n_slots = ss[i]->n_slots, st.n_instr = (uint32_t)ss[i]->instr.size();
    st.n_levels = (uint32_t)ss[i]->levels.size() - 1, st.max_level_width = ss[i]->max_width;
    st.off_levels = (uint32_t)off, off = align16(off + ss[i]->levels.size() * 4);
    st.off_instr = (uint32_t)off, off = align16(off + ss[i]->instr.size() * sizeof(FlatInstr));
  }
  std::vector<FlatStream> part_tab(parts.size());
  if (!parts.empty()) {
    h.n_parts = (uint32_t)parts.size();
    h.off_parts = (uint32_t)off, off = align16(off + parts.size() * sizeof(FlatStream));
    for (size_t q = 0; q < parts.size(); q++) {
      FlatStream &st = part_tab[q];
      st = FlatStream{};
      st.n_slots = parts[q].n_slots, st.n_instr = (uint32_t)parts[q].instr.size();
      st.n_levels = (uint32_t)parts[q].levels.size() - 1, st.max_level_width = parts[q].max_width;
      st.off_levels = (uint32_t)off, off = align16(off + parts[q].levels.size() * 4);
      st.off_instr = (uint32_t)off, off = align16(off + parts[q].instr.size() * sizeof(FlatInstr));
    }
  }
  if (rowsch.ok) {
    FlatRows &r = h.rows;
    r.n_slots = rowsch.n_slots, r.n_instr = (uint32_t)rowsch.instr.size(), r.n_segs = rowsch.n_segs, r.n_warps = rowsch.n_warps, r.lag = rowsch.lag;
    r.n_side = (uint32_t)rowsch.side.size(), r.n_from_reg = rowsch.n_from_reg, r.n_no_store = rowsch.n_no_store;
    r.off_rowtab = (uint32_t)off, off = align16(off + rowsch.rowtab.size() * 4);
    r.off_instr = (uint32_t)off, off = align16(off + rowsch.instr.size() * sizeof(FlatInstr));
    r.off_sidetab = (uint32_t)off, off = align16(off + rowsch.sidetab.size() * 4);
    r.off_side = (uint32_t)off, off = align16(off + rowsch.side.size() * sizeof(FlatInstr));
  }
  if (off >= 0xFFFFFFFFull) fail(CLG_E_UNSUPPORTED, "flat buffer exceeds 4 GiB");
  h.total_bytes = (uint32_t)off;
  f.bytes.assign(off, 0);
  std::memcpy(f.bytes.data(), &h, sizeof h);
  std::memcpy(f.bytes.data() + h.off_state_regs, state_regs.data(), state_regs.size() * 8);
  std::memcpy(f.bytes.data() + h.off_out_regs, outs.data(), outs.size() * 8);
  for (int i = 0; i < 2; i++) {
    if (!ss[i]->ok) continue;
    std::memcpy(f.bytes.data() + h.stream[i].off_levels, ss[i]->levels.data(), ss[i]->levels.size() * 4);
    std::memcpy(f.bytes.data() + h.stream[i].off_instr, ss[i]->instr.data(), ss[i]->instr.size() * sizeof(FlatInstr));
  }
  if (rowsch.ok) {
    std::memcpy(f.bytes.data() + h.rows.off_rowtab, rowsch.rowtab.data(), rowsch.rowtab.size() * 4);
    std::memcpy(f.bytes.data() + h.rows.off_instr, rowsch.instr.data(), rowsch.instr.size() * sizeof(FlatInstr));
    std::memcpy(f.bytes.data() + h.rows.off_sidetab, rowsch.sidetab.data(), rowsch.sidetab.size() * 4);
    if (!rowsch.side.empty()) std::memcpy(f.bytes.data() + h.rows.off_side, rowsch.side.data(), rowsch.side.size() * sizeof(FlatInstr));
  }
  if (!parts.empty()) std::memcpy(f.bytes.data() + h.off_parts, part_tab.data(), part_tab.size() * sizeof(FlatStream));
  for (size_t q = 0; q < parts.size(); q++) {
    std::memcpy(f.bytes.data() + part_tab[q].off_levels, parts[q].levels.data(), parts[q].levels.size() * 4);
    std::memcpy(f.bytes.data() + part_tab[q].off_instr, parts[q].instr.data(), parts[q].instr.size() * sizeof(FlatInstr));
  }
  return f;
}

// Everything a kernel relies on is checked here once, so the device code carries no bounds checks:
// index ranges, definition-before-use, and the hazards each execution order must be free of.
void validate_flat(const Flat &f) {
  auto bad = [](const std::string &m) { fail(CLG_E_FORMAT, "flat buffer: " + m); };
  if (f.bytes.size() < sizeof(FlatHeader)) bad("truncated header");
  const FlatHeader &h = f.header();
  if (h.magic != FLAT_MAGIC) bad("bad magic");
  if (h.version != FLAT_VERSION) bad("unsupported version");
  if (h.total_bytes != f.bytes.size()) bad("size mismatch");
  auto span_ok = [&](uint32_t off, uint64_t bytes) { return off % 16 == 0 && (uint64_t)off + bytes <= f.bytes.size(); };
  if (!span_ok(h.off_state_regs, (uint64_t)h.n_state * 8) || !span_ok(h.off_out_regs, (uint64_t)h.n_outputs * 8))
    bad("register tables out of range");
  const uint32_t n_in_rows = h.n_inputs + h.n_state, n_out_rows = h.n_outputs + h.n_state;
  if (h.n_parts && !span_ok(h.off_parts, (uint64_t)h.n_parts * sizeof(FlatStream))) bad("part table out of range");
  // all output/state rows stored exactly... at most once across the parts of the partitioned form
  std::vector<uint8_t> row_in_part(h.n_parts ? n_out_rows : 0, 0);
  std::vector<uint32_t> state_load_part(h.n_parts ? h.n_state : 0, NONE);  // part that loads each carried row
  for (uint32_t q = 0; q < h.n_parts; q++) {
    const FlatStream &st = f.parts()[q];
    if (st.n_instr == 0) continue;
    if (st.off_instr % 16 != 0 || (uint64_t)st.off_instr + (uint64_t)st.n_instr * 16 > f.bytes.size()) bad("stream out of range");
    const FlatInstr *in = reinterpret_cast<const FlatInstr *>(f.bytes.data() + st.off_instr);
    for (uint32_t i = 0; i < st.n_instr; i++)
      if (in[i].kind == K_LDIN && in[i].row >= h.n_inputs && in[i].row < n_in_rows) state_load_part[in[i].row - h.n_inputs] = 2 + q;
  }
  for (uint32_t sidx = 0; sidx < 2 + h.n_parts; sidx++) {
    const int s = sidx < 2 ? (int)sidx : (int)STREAM_LEVEL;
    const FlatStream &st = sidx < 2 ? h.stream[sidx] : f.parts()[sidx - 2];
    if (st.n_instr == 0) continue;
    if (!span_ok(st.off_levels, ((uint64_t)st.n_levels + 1) * 4) || !span_ok(st.off_instr, (uint64_t)st.n_instr * 16))
      bad("stream out of range");
    if (st.n_slots > 65536) bad("too many slots");
    const uint32_t *lv = reinterpret_cast<const uint32_t *>(f.bytes.data() + st.off_levels);
    const FlatInstr *in = reinterpret_cast<const FlatInstr *>(f.bytes.data() + st.off_instr);
    if (sidx >= 2)
      for (uint32_t i = 0; i < st.n_instr; i++)
        if ((in[i].kind == K_STOUT || in[i].kind == K_STCONST) && in[i].row < n_out_rows && row_in_part[in[i].row]++)
          bad("an output row is stored by two parts");
    if (st.n_levels == 0 || lv[0] != 0 || lv[st.n_levels] != st.n_instr) bad("level table does not cover the stream");
    // a level is one hazard window; the lane stream is sequential, i.e. every instruction is its own window
    const bool seq = (s == STREAM_LANE);
    std::vector<uint8_t> defined(st.n_slots, 0);
    std::vector<uint32_t> written_in(st.n_slots, NONE);
    std::vector<uint32_t> loaded_in(h.n_state, NONE), stored_in(h.n_state, NONE);
    std::vector<uint8_t> row_stored(n_out_rows, 0);  // every output / state row is stored at most once per stream
    uint32_t window = 0;
    for (uint32_t l = 0; l < st.n_levels; l++) {
      if (lv[l] > lv[l + 1] || lv[l + 1] > st.n_instr) bad("level table not monotone");
      if (lv[l + 1] - lv[l] > st.max_level_width) bad("max_level_width too small");
      for (uint32_t pass = 0; pass < (seq ? 1u : 2u); pass++)
        for (uint32_t i = lv[l]; i < lv[l + 1]; i++) {
          const FlatInstr &x = in[i];
          if (seq) window = i;
          else window = l;
          const bool reads = seq || pass == 0, writes = seq || pass == 1;
          auto rd = [&](uint16_t sl) {
            if (sl >= st.n_slots) bad("slot out of range");
            if (!defined[sl]) bad("slot read before it is written");
          };
          if (x.kind > K_STCONST) bad("unknown instruction kind");
          if (x.flags & (FLAG_FROM_REG | FLAG_NO_STORE)) bad("register forwarding outside the row stream");
          if (reads) {
            if (x.kind == K_LUT) {
              const int nf = lut_nfan(x);
              if (nf < 1) bad("LUT without operands");
              // operands the LUT does not read alias one it does (the encoders and the PTX emitter touch all three)
              rd(x.a), rd(x.b), rd(x.c);
            }
            if (x.kind == K_STOUT) rd(x.a);
            if (x.kind == K_LDIN) {
              if (x.row >= n_in_rows) bad("input row out of range");
              if (x.row >= h.n_inputs) {
                uint32_t k = x.row - h.n_inputs;
                if (stored_in[k] != NONE) bad("state row loaded after it was stored");
                loaded_in[k] = window;
              }
            }
            if (x.kind == K_STOUT || x.kind == K_STCONST) {
              if (x.row >= n_out_rows) bad("output row out of range");
              if (row_stored[x.row]++) bad("an output or state row is stored twice");
              if (x.row >= h.n_outputs) {
                uint32_t k = x.row - h.n_outputs;
                if (!seq && loaded_in[k] == window) bad("state row loaded and stored in the same level");
                if (sidx >= 2 && state_load_part[k] != NONE && state_load_part[k] != sidx)
                  bad("a state row is loaded by one part and stored by another");
              }
            }
          }
          if (writes) {
            if (x.kind == K_LUT || x.kind == K_LDIN) {
              if (x.dst >= st.n_slots) bad("slot out of range");
              if (!seq && written_in[x.dst] == window) bad("slot written twice in one level");
              written_in[x.dst] = window;
              defined[x.dst] = 1;
            }
            if ((x.kind == K_STOUT || x.kind == K_STCONST) && x.row >= h.n_outputs) stored_in[x.row - h.n_outputs] = window;
          }
        }
      if (!seq) {
        // nothing read in this level may also be written in it (other threads are still reading)
        for (uint32_t i = lv[l]; i < lv[l + 1]; i++) {
          const FlatInstr &x = in[i];
          auto chk = [&](uint16_t sl) {
            if (written_in[sl] == l) bad("slot read and written in the same level");
          };
          if (x.kind == K_LUT) chk(x.a), chk(x.b), chk(x.c);
          if (x.kind == K_STOUT) chk(x.a);
        }
      }
    }
  }

  // ---- row stream: index ranges, and the hazards of its relaxed order. A warp may run up to `lag` segments ahead of
  // the slowest one, so two accesses to a slot are only ordered if their segments are at least `lag` apart.
  const FlatRows &r = h.rows;
  if (r.n_instr == 0) return;
  if (r.lag < 1 || r.lag > 2) bad("row stream: unsupported lag");
  if (r.n_warps < 1 || r.n_warps > 32 || r.n_segs == 0 || r.n_instr % 32) bad("row stream: bad geometry");
  if (r.n_slots > 16383) bad("row stream: too many slots");
  if (!span_ok(r.off_rowtab, ((uint64_t)r.n_segs * r.n_warps + 1) * 4) || !span_ok(r.off_instr, (uint64_t)r.n_instr * 16) ||
      !span_ok(r.off_sidetab, ((uint64_t)r.n_segs + 1) * 4) || !span_ok(r.off_side, (uint64_t)r.n_side * 16))
    bad("row stream out of range");
  const uint32_t *rt = f.row_tab(), *stab = f.row_sidetab();
  const FlatInstr *ri = f.row_instrs(), *rside = f.row_side();
  if (rt[0] != 0 || rt[(size_t)r.n_segs * r.n_warps] != r.n_instr || stab[0] != 0 || stab[r.n_segs] != r.n_side) bad("row stream: tables do not cover the stream");
  constexpr uint32_t NEVER = 0x7FFFFFFFu;
  std::vector<uint32_t> wr_seg(r.n_slots, NEVER), rd_seg(r.n_slots, NEVER);  // last segment that wrote / read each slot
  std::vector<uint8_t> lane_has(r.n_warps * 32u, 0), rrow_stored(n_out_rows, 0), rstate_loaded(h.n_state, 0);
  std::vector<uint32_t> rstate_load_seg(h.n_state, NEVER);
  for (uint32_t sgm = 0; sgm < r.n_segs; sgm++) {
    // reads first, then writes, over everything of the segment: within a segment nothing is ordered between warps
    for (int pass = 0; pass < 2; pass++) {
      auto rd = [&](uint16_t sl) {
        if (sl >= r.n_slots) bad("row stream: slot out of range");
        if (wr_seg[sl] == NEVER || wr_seg[sl] + r.lag > sgm) bad("row stream: slot read before its value is visible");
        rd_seg[sl] = sgm;
      };
      auto wr = [&](uint16_t sl) {
        if (sl >= r.n_slots) bad("row stream: slot out of range");
        if (wr_seg[sl] != NEVER && wr_seg[sl] + r.lag > sgm) bad("row stream: slot written twice within the lag");
        if (rd_seg[sl] != NEVER && rd_seg[sl] + r.lag > sgm) bad("row stream: slot overwritten while it may still be read");
        wr_seg[sl] = sgm;
      };
      for (uint32_t w = 0; w < r.n_warps; w++) {
        const uint32_t b = rt[(size_t)sgm * r.n_warps + w], e = rt[(size_t)sgm * r.n_warps + w + 1];
        if (b > e || e > r.n_instr || b % 32 || e % 32) bad("row stream: row table not monotone");
        if (pass == 0)
          for (uint32_t r0 = b; r0 < e; r0 += 32) {  // the kernel dispatches at most three tables per row
            uint8_t seen[3];
            uint32_t n_seen = 0;
            for (uint32_t i = r0; i < r0 + 32; i++) {
              if (ri[i].kind != K_LUT) continue;
              uint32_t k = 0;
              while (k < n_seen && seen[k] != ri[i].tt) k++;
              if (k == n_seen) {
                if (n_seen == 3) bad("row stream: more than three truth tables in one row");
                seen[n_seen++] = ri[i].tt;
              }
            }
          }
        for (uint32_t i = b; i < e; i++) {
          const FlatInstr &x = ri[i];
          if (x.kind == K_NOP) {
            // the row's LOP3s run in every lane: a lane without an instruction loses its register (unless the whole row is empty)
            if (pass == 0) {
              bool any = false;
              for (uint32_t j = b + (i - b) / 32 * 32; j < b + (i - b) / 32 * 32 + 32 && !any; j++) any = ri[j].kind == K_LUT;
              if (any) lane_has[w * 32u + (i - b) % 32] = 0;
            }
            continue;
          }
          if (x.kind != K_LUT) bad("row stream: a row holds LUTs only");
          const int nf = lut_nfan(x);
          if (nf < 1) bad("LUT without operands");
          if (pass == 0) {
            if (x.flags & FLAG_FROM_REG) {
              if (!lane_has[w * 32u + (i - b) % 32]) bad("row stream: register operand before the lane computed anything");
            } else
              rd(x.a);
            if (nf >= 2) rd(x.b);
            if (nf >= 3) rd(x.c);
            lane_has[w * 32u + (i - b) % 32] = 1;  // (a warp's rows run one after the other)
          } else {
            if (!(x.flags & FLAG_NO_STORE)) wr(x.dst);
          }
        }
      }
      if (stab[sgm] > stab[sgm + 1] || stab[sgm + 1] > r.n_side) bad("row stream: side table not monotone");
      for (uint32_t i = stab[sgm]; i < stab[sgm + 1]; i++) {
        const FlatInstr &x = rside[i];
        if (x.kind > K_STCONST) bad("row stream: unknown side instruction");
        if (x.flags & (FLAG_FROM_REG | FLAG_NO_STORE)) bad("row stream: register forwarding in the side list");
        if (pass == 0) {
          if (x.kind == K_LUT) rd(x.a), rd(x.b), rd(x.c);
          if (x.kind == K_STOUT) rd(x.a);
          if (x.kind == K_LDIN) {
            if (x.row >= n_in_rows) bad("input row out of range");
            if (x.row >= h.n_inputs) rstate_load_seg[x.row - h.n_inputs] = sgm;
          }
          if (x.kind == K_STOUT || x.kind == K_STCONST) {
            if (x.row >= n_out_rows) bad("output row out of range");
            if (rrow_stored[x.row]++) bad("an output or state row is stored twice");
          }
        } else {
          if (x.kind == K_LUT || x.kind == K_LDIN) wr(x.dst);
        }
      }
    }
    // in-place state rows: a row's store must be ordered after its load
    for (uint32_t i = stab[sgm]; i < stab[sgm + 1]; i++) {
      const FlatInstr &x = rside[i];
      if ((x.kind == K_STOUT || x.kind == K_STCONST) && x.row >= h.n_outputs) {
        const uint32_t ls = rstate_load_seg[x.row - h.n_outputs];
        if (ls != NEVER && ls + r.lag > sgm) bad("row stream: state row stored before its load is ordered");
      }
    }
  }
  // a state row that is stored must not be loaded later
  (void)rstate_loaded;
}

}  // namespace clg
