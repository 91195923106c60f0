/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the complogic NAND-VM hot path.
 *
 * A plain-C restatement of the reference algorithm (Vandesm14/complogic). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * file's shared object; the product (complogic_b200/) never links, imports or calls it.
 *
 * Parity status: the reference is Rust and no cargo/rustc exists in the build image, so the
 * reference itself cannot be executed here. This restatement is PINNED against every known-
 * answer test the reference holds for the path (14 unit tests: simulation.rs:42-94,
 * compile.rs:211-294, gates.rs:464-645) and against the committed graph.dot artefact
 * (graph.dot:1-37) -- see tests/test_oracle_golden.py.
 *
 * What follows what:
 *   orc_run                 complogic/src/simulation.rs:16-30   (Simulation::run)
 *   orc_op (32 bytes)       complogic/src/compile.rs:7-14       (enum Op)
 *   orc_gate_create         complogic/src/gates.rs:192-456      (Gate::create)
 *   orc_compile             complogic/src/compile.rs:93-202     (Compiler::compile)
 *   orc_run_batch_*         batch semantics defined by SURVEY.md 8b: lane v == fresh all-false
 *                           Simulation followed by one run(imm_v) (or `steps` runs, state carried)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#include <pthread.h>

#define ORC_NAND 0u
#define ORC_SET 1u

/* Same footprint as Rust's `enum Op` on x86-64: 8-byte tag + three usize (compile.rs:7-14).
 * Nand(a, b, out) -> {ORC_NAND, a, b, out};  Set(id, val) -> {ORC_SET, id, val, 0}. */
typedef struct {
  uint64_t tag;
  uint64_t a, b, c;
} orc_op;

/* ------------------------------------------------------------------------------------------
 * Simulation::run -- simulation.rs:16-30.  Returns 0, or -1 where Rust would panic on an
 * out-of-bounds register index (simulation.rs:20,21,23,26).
 * ---------------------------------------------------------------------------------------- */
int orc_run(const orc_op *ops, size_t n_ops, uint8_t *registers, size_t n_regs,
            const uint8_t *immediates, size_t n_imm) {
  for (size_t i = 0; i < n_ops; i++) {
    const orc_op *op = &ops[i];
    if (op->tag == ORC_NAND) {
      if (op->a >= n_regs || op->b >= n_regs || op->c >= n_regs) return -1;
      uint8_t a = registers[op->a];
      uint8_t b = registers[op->b];
      registers[op->c] = !(a && b);
    } else {
      if (op->a >= n_regs) return -1;
      /* immediates.get(id).copied().unwrap_or(val) -- simulation.rs:26 */
      registers[op->a] = (op->a < n_imm) ? (immediates[op->a] != 0) : (uint8_t)(op->b != 0);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Gate::create -- gates.rs:192-456.  `f` holds the gate struct's fields in declaration order
 * (gates.rs:5-100); `inc` is Incrementer.val (compile.rs:19-49).  Scratch registers are drawn
 * when the child structs are BUILT, i.e. before any recursive create() call, which fixes the
 * numbering (e.g. FullAdder takes h1.s, h1.c, h2.c first: gates.rs:392-403).
 * ---------------------------------------------------------------------------------------- */
enum {
  G_NAND = 0, G_NOT, G_AND, G_OR, G_NOR, G_XOR, G_RSLATCH, G_RSLATCHTEST, G_DLATCH,
  G_HALFADDER, G_FULLADDER, G_FOURBITADDER
};

typedef struct {
  orc_op *v;
  size_t n, cap;
} opvec;

static void push(opvec *o, uint64_t tag, uint64_t a, uint64_t b, uint64_t c) {
  if (o->n == o->cap) {
    o->cap = o->cap ? o->cap * 2 : 64;
    o->v = (orc_op *)realloc(o->v, o->cap * sizeof(orc_op));
  }
  orc_op op = {tag, a, b, c};
  o->v[o->n++] = op;
}

static uint64_t nxt(uint64_t *inc) { return (*inc)++; } /* Incrementer::next, compile.rs:33-38 */

static void create(int kind, const uint64_t *f, uint64_t *inc, opvec *o) {
  switch (kind) {
  case G_NAND: /* gates.rs:195-197 */
    push(o, ORC_NAND, f[0], f[1], f[2]);
    break;
  case G_NOT: /* gates.rs:198-200 */
    push(o, ORC_NAND, f[0], f[0], f[1]);
    break;
  case G_AND: { /* gates.rs:201-217 */
    uint64_t nand[3] = {f[0], f[1], nxt(inc)};
    uint64_t not_[2] = {nand[2], f[2]};
    create(G_NAND, nand, inc, o);
    create(G_NOT, not_, inc, o);
    break;
  }
  case G_OR: { /* gates.rs:218-241 */
    uint64_t na[3] = {f[0], f[0], nxt(inc)};
    uint64_t nb[3] = {f[1], f[1], nxt(inc)};
    uint64_t n[3] = {na[2], nb[2], f[2]};
    create(G_NAND, na, inc, o);
    create(G_NAND, nb, inc, o);
    create(G_NAND, n, inc, o);
    break;
  }
  case G_NOR: { /* gates.rs:242-258 */
    uint64_t or_[3] = {f[0], f[1], nxt(inc)};
    uint64_t not_[2] = {or_[2], f[2]};
    create(G_OR, or_, inc, o);
    create(G_NOT, not_, inc, o);
    break;
  }
  case G_XOR: { /* gates.rs:259-282 */
    uint64_t or_[3] = {f[0], f[1], nxt(inc)};
    uint64_t nand[3] = {f[0], f[1], nxt(inc)};
    uint64_t and_[3] = {or_[2], nand[2], f[2]};
    create(G_OR, or_, inc, o);
    create(G_NAND, nand, inc, o);
    create(G_AND, and_, inc, o);
    break;
  }
  case G_RSLATCH:       /* gates.rs:283-310 */
  case G_RSLATCHTEST: { /* gates.rs:312-339 (test-only variant: same parts, reversed emission) */
    uint64_t q_patch = nxt(inc);
    uint64_t qn_patch = nxt(inc);
    uint64_t nor1[3] = {f[0], qn_patch, q_patch};
    uint64_t nor2[3] = {f[1], q_patch, qn_patch};
    uint64_t orq[3] = {qn_patch, qn_patch, f[2]};
    if (kind == G_RSLATCH) {
      create(G_NOR, nor1, inc, o);
      create(G_NOR, nor2, inc, o);
      create(G_OR, orq, inc, o);
    } else {
      create(G_OR, orq, inc, o);
      create(G_NOR, nor2, inc, o);
      create(G_NOR, nor1, inc, o);
    }
    break;
  }
  case G_DLATCH: { /* gates.rs:340-372; f = {d, e, q} */
    uint64_t not_[2] = {f[0], nxt(inc)};
    uint64_t and1[3] = {not_[1], f[1], nxt(inc)};
    uint64_t and2[3] = {f[1], f[0], nxt(inc)};
    uint64_t rs[3] = {and2[2], and1[2], f[2]}; /* s/r deliberately swapped: gates.rs:357-363 */
    create(G_NOT, not_, inc, o);
    create(G_AND, and1, inc, o);
    create(G_AND, and2, inc, o);
    create(G_RSLATCH, rs, inc, o);
    break;
  }
  case G_HALFADDER: { /* gates.rs:373-390; f = {a, b, s, c} */
    uint64_t xor_[3] = {f[0], f[1], f[2]};
    uint64_t and_[3] = {f[0], f[1], f[3]};
    create(G_XOR, xor_, inc, o);
    create(G_AND, and_, inc, o);
    break;
  }
  case G_FULLADDER: { /* gates.rs:391-416; f = {a, b, cin, s, cout} */
    uint64_t h1[4] = {f[0], f[1], 0, 0};
    h1[2] = nxt(inc);
    h1[3] = nxt(inc);
    uint64_t h2[4] = {h1[2], f[2], f[3], 0};
    h2[3] = nxt(inc);
    uint64_t or_[3] = {h1[3], h2[3], f[4]};
    create(G_HALFADDER, h1, inc, o);
    create(G_HALFADDER, h2, inc, o);
    create(G_OR, or_, inc, o);
    break;
  }
  case G_FOURBITADDER: { /* gates.rs:417-454; f = {a1..a4, b1..b4, s1..s4, cout} */
    uint64_t fa1[5] = {f[0], f[4], 0, f[8], 0};
    fa1[2] = nxt(inc); /* cin: a register nobody writes (gates.rs:421) */
    fa1[4] = nxt(inc);
    uint64_t fa2[5] = {f[1], f[5], fa1[4], f[9], 0};
    fa2[4] = nxt(inc);
    uint64_t fa3[5] = {f[2], f[6], fa2[4], f[10], 0};
    fa3[4] = nxt(inc);
    uint64_t fa4[5] = {f[3], f[7], fa3[4], f[11], f[12]};
    create(G_FULLADDER, fa1, inc, o);
    create(G_FULLADDER, fa2, inc, o);
    create(G_FULLADDER, fa3, inc, o);
    create(G_FULLADDER, fa4, inc, o);
    break;
  }
  default:
    break;
  }
}

/* Lower one gate. Returns the op count and hands back a malloc'd array the caller frees with
 * orc_free. *inc is advanced exactly as Gate::create advances its Incrementer. */
size_t orc_gate_create(int kind, const uint64_t *fields, uint64_t *inc, orc_op **out) {
  opvec o = {0, 0, 0};
  create(kind, fields, inc, &o);
  *out = o.v;
  return o.n;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Compiler::compile -- compile.rs:93-202.  petgraph's DiGraph is restated as the only things
 * the reference uses of it: a node-weight array, an edge list in insertion order (what Dot
 * prints, compile.rs:150) and in/out neighbour walks (compile.rs:162-163, 177-181).
 *
 * Returns 0 on success; -1 where the reference would panic at compile.rs:137/141 (graph index
 * out of bounds, e.g. FourBitAdder's unwritten cin).  dot_out (optional) receives the text of
 * the Dot dump the reference writes to ./graph.dot; this restatement never touches the disk.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t *v;
  size_t n, cap;
} u64vec;

static void upush(u64vec *q, uint64_t x) {
  if (q->n == q->cap) {
    q->cap = q->cap ? q->cap * 2 : 64;
    q->v = (uint64_t *)realloc(q->v, q->cap * sizeof(uint64_t));
  }
  q->v[q->n++] = x;
}

static int cmp_u64(const void *x, const void *y) {
  uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
  return a < b ? -1 : a > b;
}

static void fmt_op(char *buf, size_t cap, const orc_op *op) {
  if (op->tag == ORC_NAND)
    snprintf(buf, cap, "Nand(%llu, %llu, %llu)", (unsigned long long)op->a,
             (unsigned long long)op->b, (unsigned long long)op->c);
  else
    snprintf(buf, cap, "Set(%llu, %s)", (unsigned long long)op->a, op->b ? "true" : "false");
}

int orc_compile(uint64_t immediate_count, uint64_t incrementer_val, const int *gate_kinds,
                const uint64_t *gate_fields /* 13 per gate */, size_t n_gates, orc_op **ops_out,
                size_t *n_ops_out, uint64_t *n_registers_out, char **dot_out) {
  *ops_out = NULL;
  *n_ops_out = 0;
  if (dot_out) *dot_out = NULL;

  /* reset_ops: compile.rs:74-80, :94 */
  opvec sops = {0, 0, 0};
  for (uint64_t i = 0; i < immediate_count; i++) push(&sops, ORC_SET, i, 0, 0);

  if (n_gates == 0) { /* compile.rs:96-101 */
    free(sops.v);
    *n_registers_out = immediate_count;
    return 0;
  }

  uint64_t inc = incrementer_val; /* clone: compile.rs:105 */
  for (size_t g = 0; g < n_gates; g++) create(gate_kinds[g], gate_fields + 13 * g, &inc, &sops);

  size_t nn = sops.n; /* node_count */
  orc_op *weight = (orc_op *)malloc((nn ? nn : 1) * sizeof(orc_op));
  memcpy(weight, sops.v, nn * sizeof(orc_op));

  /* edges: compile.rs:115-125 */
  u64vec esrc = {0, 0, 0}, edst = {0, 0, 0};
  for (size_t i = 0; i < nn; i++) {
    const orc_op *op = &sops.v[i];
    if (op->tag == ORC_NAND && op->a < nn && op->b < nn && op->c < nn) {
      upush(&esrc, op->a), upush(&edst, op->c);
      upush(&esrc, op->b), upush(&edst, op->c);
    }
  }
  size_t ne = esrc.n;

  /* relabel + seed: compile.rs:133-148 */
  uint8_t *todo = (uint8_t *)calloc(inc + nn + 1, 1); /* nodes_to_process as a bitmap */
  u64vec queue = {0, 0, 0}, next = {0, 0, 0};
  int rc = 0;
  for (size_t i = 0; i < nn && rc == 0; i++) {
    orc_op op = sops.v[i];
    if (op.tag == ORC_NAND) {
      if (op.c >= nn) { rc = -1; break; }
      weight[op.c] = op;
      todo[op.c] = 1;
    } else {
      if (op.a >= nn) { rc = -1; break; }
      weight[op.a] = op;
      upush(&queue, op.a);
      todo[op.a] = 1;
    }
  }
  if (rc != 0) goto done;

  if (dot_out) { /* Dot::with_config(&graph, &[]) via {:?}: compile.rs:150 */
    size_t cap = 64 + nn * 64 + ne * 48, len = 0;
    char *d = (char *)malloc(cap), lbl[96];
    len += (size_t)snprintf(d + len, cap - len, "digraph {\n");
    for (size_t i = 0; i < nn; i++) {
      fmt_op(lbl, sizeof lbl, &weight[i]);
      len += (size_t)snprintf(d + len, cap - len, "    %zu [ label = \"%s\" ]\n", i, lbl);
    }
    for (size_t e = 0; e < ne; e++)
      len += (size_t)snprintf(d + len, cap - len, "    %llu -> %llu [ label = \"()\" ]\n",
                              (unsigned long long)esrc.v[e], (unsigned long long)edst.v[e]);
    len += (size_t)snprintf(d + len, cap - len, "}\n");
    *dot_out = d;
  }

  {
    /* adjacency (CSR) for the neighbour walks */
    size_t *in_off = (size_t *)calloc(nn + 2, sizeof(size_t));
    size_t *out_off = (size_t *)calloc(nn + 2, sizeof(size_t));
    for (size_t e = 0; e < ne; e++) in_off[edst.v[e] + 1]++, out_off[esrc.v[e] + 1]++;
    for (size_t i = 0; i < nn; i++) in_off[i + 1] += in_off[i], out_off[i + 1] += out_off[i];
    uint64_t *in_adj = (uint64_t *)malloc((ne ? ne : 1) * sizeof(uint64_t));
    uint64_t *out_adj = (uint64_t *)malloc((ne ? ne : 1) * sizeof(uint64_t));
    size_t *ic = (size_t *)calloc(nn + 1, sizeof(size_t)), *oc = (size_t *)calloc(nn + 1, sizeof(size_t));
    for (size_t e = 0; e < ne; e++) {
      in_adj[in_off[edst.v[e]] + ic[edst.v[e]]++] = esrc.v[e];
      out_adj[out_off[esrc.v[e]] + oc[esrc.v[e]]++] = edst.v[e];
    }
    free(ic), free(oc);

    opvec out = {0, 0, 0};
    int recursion_flag = 0; /* compile.rs:154 */
    for (;;) {              /* compile.rs:155-196 */
      for (size_t qi = 0; qi < queue.n; qi++) {
        uint64_t node = queue.v[qi];
        if (!todo[node]) continue;
        int requires_new_layer = 0;
        if (!recursion_flag)
          for (size_t k = in_off[node]; k < in_off[node + 1]; k++)
            if (todo[in_adj[k]]) { requires_new_layer = 1; break; }
        if (requires_new_layer) {
          upush(&next, node);
        } else {
          push(&out, weight[node].tag, weight[node].a, weight[node].b, weight[node].c);
          todo[node] = 0;
          for (size_t k = out_off[node]; k < out_off[node + 1]; k++) upush(&next, out_adj[k]);
        }
      }
      qsort(queue.v, queue.n, sizeof(uint64_t), cmp_u64);
      qsort(next.v, next.n, sizeof(uint64_t), cmp_u64);
      recursion_flag = queue.n == next.n &&
                       (queue.n == 0 || memcmp(queue.v, next.v, queue.n * sizeof(uint64_t)) == 0);
      u64vec t = queue;
      queue = next;
      next = t;
      next.n = 0;
      if (queue.n == 0) break;
    }
    *ops_out = out.v;
    *n_ops_out = out.n;
    *n_registers_out = inc; /* vec![false; incrementer.val]: compile.rs:198-201 */
    free(in_off), free(out_off), free(in_adj), free(out_adj);
  }

done:
  free(sops.v), free(weight), free(esrc.v), free(edst.v), free(todo), free(queue.v), free(next.v);
  if (rc != 0 && dot_out && *dot_out) { free(*dot_out); *dot_out = NULL; }
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers.  Bit-sliced layout (the device ABI): row r, word w holds
 * vectors 32w..32w+31 of immediate/output r; vector v <-> bit v%32 of word v/32.
 * Lane v == Simulation{ops, registers: all false}; run(imm_v) `steps` times.
 * imm is [steps][n_imm][stride] (a fresh immediates slice per step), out is the value of
 * out_regs after the LAST step unless all_steps, in which case out is [steps][n_out][stride].
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const orc_op *ops;
  size_t n_ops, n_regs;
  const uint32_t *imm;
  size_t n_imm, stride, steps;
  const uint64_t *out_regs;
  size_t n_out;
  uint32_t *out;
  int all_steps;
  size_t v0, v1;
  int rc;
} batch_job;

static void *scalar_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  uint8_t *regs = (uint8_t *)malloc(j->n_regs ? j->n_regs : 1);
  uint8_t *imm = (uint8_t *)malloc(j->n_imm ? j->n_imm : 1);
  j->rc = 0;
  for (size_t v = j->v0; v < j->v1; v++) {
    memset(regs, 0, j->n_regs);
    size_t w = v >> 5, bit = v & 31;
    for (size_t s = 0; s < j->steps; s++) {
      const uint32_t *im = j->imm + s * j->n_imm * j->stride;
      for (size_t i = 0; i < j->n_imm; i++) imm[i] = (im[i * j->stride + w] >> bit) & 1u;
      if (orc_run(j->ops, j->n_ops, regs, j->n_regs, imm, j->n_imm) != 0) { j->rc = -1; break; }
      if (j->all_steps || s + 1 == j->steps) {
        uint32_t *o = j->out + (j->all_steps ? s * j->n_out * j->stride : 0);
        for (size_t k = 0; k < j->n_out; k++) {
          if (j->out_regs[k] >= j->n_regs) { j->rc = -1; break; }
          uint32_t *p = &o[k * j->stride + w];
          uint32_t m = 1u << bit;
          /* workers own disjoint word ranges, so plain read-modify-write is race free */
          *p = regs[j->out_regs[k]] ? (*p | m) : (*p & ~m);
        }
      }
    }
    if (j->rc) break;
  }
  free(regs), free(imm);
  return NULL;
}

/* Scalar (one vector at a time, byte registers, 32-byte ops) -- the literal reference shape.
 * threads>1 splits the vectors into word-aligned contiguous ranges, one pthread each. */
int orc_run_batch_scalar(const orc_op *ops, size_t n_ops, size_t n_regs, const uint32_t *imm,
                         size_t n_imm, size_t stride, size_t n_vectors, size_t steps,
                         const uint64_t *out_regs, size_t n_out, uint32_t *out, int all_steps,
                         int threads) {
  if (threads < 1) threads = 1;
  size_t words = (n_vectors + 31) / 32;
  if ((size_t)threads > words) threads = words ? (int)words : 1;
  batch_job *jobs = (batch_job *)calloc((size_t)threads, sizeof(batch_job));
  pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; t++) {
    size_t w0 = words * (size_t)t / (size_t)threads, w1 = words * (size_t)(t + 1) / (size_t)threads;
    batch_job j = {ops, n_ops, n_regs, imm, n_imm, stride, steps, out_regs, n_out, out, all_steps,
                   w0 * 32, w1 * 32 < n_vectors ? w1 * 32 : n_vectors, 0};
    jobs[t] = j;
    if (threads == 1) scalar_worker(&jobs[t]);
    else pthread_create(&tid[t], NULL, scalar_worker, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < threads; t++) {
    if (threads > 1) pthread_join(tid[t], NULL);
    if (jobs[t].rc) rc = -1;
  }
  free(jobs), free(tid);
  return rc;
}

/* Bit-sliced CPU evaluator (oracle tier L1): same op list, same order, but every register is a
 * 64-bit word carrying 64 lanes; !(a && b) becomes ~(a & b).  Validated against the scalar
 * driver in tests; used where the scalar one would take minutes. */
int orc_run_batch_sliced(const orc_op *ops, size_t n_ops, size_t n_regs, const uint32_t *imm,
                         size_t n_imm, size_t stride, size_t n_vectors, size_t steps,
                         const uint64_t *out_regs, size_t n_out, uint32_t *out, int all_steps) {
  for (size_t i = 0; i < n_ops; i++) {
    if (ops[i].tag == ORC_NAND) {
      if (ops[i].a >= n_regs || ops[i].b >= n_regs || ops[i].c >= n_regs) return -1;
    } else if (ops[i].a >= n_regs) return -1;
  }
  for (size_t k = 0; k < n_out; k++)
    if (out_regs[k] >= n_regs) return -1;
  size_t words = (n_vectors + 31) / 32;
  uint64_t *regs = (uint64_t *)malloc((n_regs ? n_regs : 1) * sizeof(uint64_t));
  for (size_t w = 0; w < words; w += 2) {
    int two = (w + 1 < words);
    memset(regs, 0, n_regs * sizeof(uint64_t));
    for (size_t s = 0; s < steps; s++) {
      const uint32_t *im = imm + s * n_imm * stride;
      for (size_t i = 0; i < n_ops; i++) {
        const orc_op *op = &ops[i];
        if (op->tag == ORC_NAND) {
          regs[op->c] = ~(regs[op->a] & regs[op->b]);
        } else if (op->a < n_imm) {
          uint64_t lo = im[op->a * stride + w], hi = two ? im[op->a * stride + w + 1] : 0;
          regs[op->a] = lo | (hi << 32);
        } else {
          regs[op->a] = op->b ? ~0ull : 0ull;
        }
      }
      if (all_steps || s + 1 == steps) {
        uint32_t *o = out + (all_steps ? s * n_out * stride : 0);
        for (size_t k = 0; k < n_out; k++) {
          uint64_t r = regs[out_regs[k]];
          o[k * stride + w] = (uint32_t)r;
          if (two) o[k * stride + w + 1] = (uint32_t)(r >> 32);
        }
      }
    }
  }
  free(regs);
  /* clear lanes past n_vectors in the last word so results compare word-for-word */
  if (n_vectors & 31) {
    uint32_t m = (1u << (n_vectors & 31)) - 1u;
    size_t blocks = all_steps ? steps : 1;
    for (size_t s = 0; s < blocks; s++)
      for (size_t k = 0; k < n_out; k++) out[(s * n_out + k) * stride + words - 1] &= m;
  }
  return 0;
}

double orc_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
