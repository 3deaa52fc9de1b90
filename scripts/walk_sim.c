// walk_sim.c -- CPU model of the tree-walk variants, used to choose the group-walk design before
// spending GPU time.  Builds the reference-shaped octree (one particle per leaf, child centre =
// centre +- size/4) over positions read from a raw float64 file, flattens it to the pre-order
// (centre, s2, skip) array the GPU walk uses, and counts for groups of G Morton-consecutive
// targets:
//   per-target accepted nodes under the reference criterion (size/|centre-x| < theta)
//   entries the per-lane warp scan steps through (union over the lanes)          [walk_kernel]
//   windows of W pre-order entries and list length of the group-MAC window scan  [walk_group_kernel]
// Build: gcc -O2 -o /tmp/walk_sim scripts/walk_sim.c -lm
// Usage: walk_sim pos.f64 N theta G W stride
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct Node {
  double c[3], size;
  int child[8];
  int particle;  // >= 0: leaf
  int count;
} Node;

static Node *nodes;
static int nnodes, capnodes;
static const double *pos;

static int new_node(const double c[3], double size) {
  if (nnodes == capnodes) {
    capnodes *= 2;
    nodes = realloc(nodes, sizeof(Node) * capnodes);
  }
  Node *n = &nodes[nnodes];
  memcpy(n->c, c, sizeof(double) * 3);
  n->size = size;
  for (int k = 0; k < 8; k++) n->child[k] = -1;
  n->particle = -1;
  n->count = 0;
  return nnodes++;
}
static int branch(const double *p, const double *c) {
  return (p[0] > c[0]) | ((p[1] > c[1]) << 1) | ((p[2] > c[2]) << 2);
}
static void insert(int ni, int pi) {
  for (;;) {
    Node *n = &nodes[ni];
    if (n->count == 0) { n->particle = pi; n->count = 1; return; }
    if (n->count == 1) {  // split leaf
      int old = n->particle;
      n->particle = -1;
      int b = branch(pos + 3 * old, n->c);
      double cc[3];
      for (int k = 0; k < 3; k++) cc[k] = n->c[k] + (((b >> k) & 1) ? 0.25 : -0.25) * n->size;
      int ch = new_node(cc, 0.5 * n->size);
      n = &nodes[ni];
      n->child[b] = ch;
      nodes[ch].particle = old;
      nodes[ch].count = 1;
    }
    n->count++;
    int b = branch(pos + 3 * pi, n->c);
    if (n->child[b] < 0) {
      double cc[3];
      for (int k = 0; k < 3; k++) cc[k] = n->c[k] + (((b >> k) & 1) ? 0.25 : -0.25) * n->size;
      int ch = new_node(cc, 0.5 * n->size);
      n = &nodes[ni];
      n->child[b] = ch;
    }
    ni = n->child[b];
  }
}

// flattened
static float (*ecen)[4];  // cx cy cz s2 (s2 < 0: leaf)
static int *eskip;
static int *eleafp;
static int nent;
static int *order;  // morton order of particles
static int nord;
static double inv_theta2;

static void flatten(int ni) {
  Node *n = &nodes[ni];
  int e = nent++;
  if (n->count == 1) {
    const double *p = pos + 3 * n->particle;
    ecen[e][0] = p[0]; ecen[e][1] = p[1]; ecen[e][2] = p[2]; ecen[e][3] = -1.f;
    eleafp[e] = n->particle;
    order[nord++] = n->particle;
    eskip[e] = e + 1;
    return;
  }
  ecen[e][0] = n->c[0]; ecen[e][1] = n->c[1]; ecen[e][2] = n->c[2];
  ecen[e][3] = (float)(n->size * n->size * inv_theta2);
  eleafp[e] = -1;
  for (int k = 0; k < 8; k++) if (n->child[k] >= 0) flatten(n->child[k]);
  eskip[e] = nent;
}

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage\n"); return 1; }
  int N = atoi(argv[2]);
  double theta = atof(argv[3]);
  int G = atoi(argv[4]), W = atoi(argv[5]), stride = atoi(argv[6]);
  double *P = malloc(sizeof(double) * 3 * N);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(P, sizeof(double) * 3, N, f) != (size_t)N) { fprintf(stderr, "read\n"); return 1; }
  fclose(f);
  pos = P;
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < N; i++) for (int k = 0; k < 3; k++) {
    if (P[3 * i + k] < mn[k]) mn[k] = P[3 * i + k];
    if (P[3 * i + k] > mx[k]) mx[k] = P[3 * i + k];
  }
  double c[3], size = 0;
  for (int k = 0; k < 3; k++) { c[k] = 0.5 * (mn[k] + mx[k]); if (mx[k] - mn[k] > size) size = mx[k] - mn[k]; }
  capnodes = 3 * N + 16;
  nodes = malloc(sizeof(Node) * capnodes);
  int root = new_node(c, size * 1.0000001);
  for (int i = 0; i < N; i++) insert(root, i);
  inv_theta2 = 1.0 / (theta * theta);
  ecen = malloc(sizeof(float) * 4 * nnodes);
  eskip = malloc(sizeof(int) * nnodes);
  eleafp = malloc(sizeof(int) * nnodes);
  order = malloc(sizeof(int) * N);
  flatten(root);
  fprintf(stderr, "N=%d nodes=%d entries=%d\n", N, nnodes, nent);

  long ngroups = 0;
  double sum_acc = 0, sum_union = 0, sum_win = 0, sum_list = 0, sum_win2 = 0, sum_tested = 0;
  double max_win = 0, max_list = 0, sum_it4 = 0, sum_t4 = 0, max_sp4 = 0;
  long *lists = malloc(sizeof(long) * (N / G + 1)), *its5 = malloc(sizeof(long) * (N / G + 1)), *sps5 = malloc(sizeof(long) * (N / G + 1));
  double sum_it5 = 0, max_sp5 = 0, sum_l6 = 0, sum_it6 = 0; long *lists6 = malloc(sizeof(long) * (N / G + 1));
  float *tx = malloc(sizeof(float) * 3 * G);
  int *until = malloc(sizeof(int) * G);
  for (int g0 = 0; g0 + G <= N; g0 += G * stride) {
    float bmn[3] = {1e30f, 1e30f, 1e30f}, bmx[3] = {-1e30f, -1e30f, -1e30f};
    for (int t = 0; t < G; t++) for (int k = 0; k < 3; k++) {
      float v = (float)P[3 * order[g0 + t] + k];
      tx[3 * t + k] = v;
      if (v < bmn[k]) bmn[k] = v;
      if (v > bmx[k]) bmx[k] = v;
    }
    float bc[3], bh[3];
    for (int k = 0; k < 3; k++) { bc[k] = 0.5f * (bmn[k] + bmx[k]); bh[k] = 0.5f * (bmx[k] - bmn[k]); }
    // (1) per-lane scan as in walk_kernel, G lanes
    for (int t = 0; t < G; t++) until[t] = 0;
    long acc = 0, un = 0;
    int i = 0;
    while (i < nent) {
      un++;
      int next = 0x7fffffff;
      for (int t = 0; t < G; t++) {
        int nx;
        if (i >= until[t]) {
          float dx = ecen[i][0] - tx[3 * t], dy = ecen[i][1] - tx[3 * t + 1], dz = ecen[i][2] - tx[3 * t + 2];
          float d2 = dx * dx + dy * dy + dz * dz;
          if (ecen[i][3] < d2) { acc++; until[t] = eskip[i]; nx = until[t]; }
          else nx = i + 1;
        } else nx = until[t];
        if (nx < next) next = nx;
      }
      i = next;
    }
    sum_acc += (double)acc / G;
    sum_union += un;
    // (2) window scan with the group criterion
    long win = 0, list = 0, tested = 0;
    i = 0;
    while (i < nent) {
      win++;
      int covered = i;  // entries < covered are dead
      int j;
      for (j = i; j < i + W && j < nent; j++) {
        tested++;
        float d2 = 0;
        for (int k = 0; k < 3; k++) {
          float d = fabsf(ecen[j][k] - bc[k]) - bh[k];
          if (d > 0) d2 += d * d;
        }
        int accept = ecen[j][3] < d2;  // leaves: -1 < d2 always
        if (j >= covered && accept) list++;
        if (accept && eskip[j] > covered) covered = eskip[j];
      }
      i = (covered > j) ? covered : j;
    }
    sum_win += win;
    sum_list += list;
    sum_tested += tested;
    if (win > max_win) max_win = win;
    if (list > max_list) max_list = list;
    // (4) stack traversal: pop up to 32 entries, test vs bbox, accept -> list, open -> push children
    {
      static int stack[1 << 16];
      int sp = 0;
      long iters = 0, tst = 0, lst = 0, maxsp = 0;
      stack[sp++] = 0;
      while (sp > 0) {
        int take = sp < 32 ? sp : 32;
        int batch[32];
        for (int t = 0; t < take; t++) batch[t] = stack[--sp];
        iters++;
        for (int t = 0; t < take; t++) {
          int j = batch[t];
          tst++;
          float d2 = 0;
          for (int k = 0; k < 3; k++) {
            float d = fabsf(ecen[j][k] - bc[k]) - bh[k];
            if (d > 0) d2 += d * d;
          }
          if (ecen[j][3] < d2) lst++;
          else for (int c2 = j + 1; c2 < eskip[j]; c2 = eskip[c2]) stack[sp++] = c2;
        }
        if (sp > maxsp) maxsp = sp;
      }
      sum_it4 += iters; sum_t4 += tst; if (maxsp > max_sp4) max_sp4 = maxsp;
      if (lst != list) { fprintf(stderr, "mismatch %ld %ld\n", lst, list); }
    }
    // (5) stack of sibling chains (first, end): pop up to 32 chains, test `first`, push the rest of
    // the chain and, if opened, the chain of its children
    {
      static int st5[1 << 16][2];
      int sp = 0;
      long iters = 0, tst = 0, lst = 0, maxsp = 0;
      st5[0][0] = 0; st5[0][1] = nent; sp = 1;
      while (sp > 0) {
        int take = sp < 32 ? sp : 32;
        int bf[32], be[32];
        for (int t = 0; t < take; t++) { --sp; bf[t] = st5[sp][0]; be[t] = st5[sp][1]; }
        iters++;
        // push order: remainders first, then children (children end up on top = depth first)
        for (int t = take - 1; t >= 0; t--) if (eskip[bf[t]] < be[t]) { st5[sp][0] = eskip[bf[t]]; st5[sp][1] = be[t]; sp++; }
        for (int t = take - 1; t >= 0; t--) {
          int j = bf[t];
          tst++;
          float d2 = 0;
          for (int k = 0; k < 3; k++) {
            float d = fabsf(ecen[j][k] - bc[k]) - bh[k];
            if (d > 0) d2 += d * d;
          }
          if (ecen[j][3] < d2) lst++;
          else { st5[sp][0] = j + 1; st5[sp][1] = eskip[j]; sp++; }
        }
        if (sp > maxsp) maxsp = sp;
      }
      sum_it5 += iters; if (maxsp > max_sp5) max_sp5 = maxsp;
      if (lst != list) { fprintf(stderr, "mismatch5 %ld %ld\n", lst, list); }
      its5[ngroups] = iters; sps5[ngroups] = maxsp;
    }
    // (6) two boxes: split the group at the largest distance jump between Morton-consecutive targets
    {
      int split = 1; float best = -1;
      for (int t = 0; t + 1 < G; t++) {
        float d = 0; for (int k = 0; k < 3; k++) { float q = tx[3 * (t + 1) + k] - tx[3 * t + k]; d += q * q; }
        if (d > best) { best = d; split = t + 1; }
      }
      float c2[2][3], h2[2][3];
      for (int b = 0; b < 2; b++) {
        float mn2[3] = {1e30f, 1e30f, 1e30f}, mx2[3] = {-1e30f, -1e30f, -1e30f};
        for (int t = (b ? split : 0); t < (b ? G : split); t++) for (int k = 0; k < 3; k++) {
          float v = tx[3 * t + k]; if (v < mn2[k]) mn2[k] = v; if (v > mx2[k]) mx2[k] = v; }
        for (int k = 0; k < 3; k++) { c2[b][k] = 0.5f * (mn2[k] + mx2[k]); h2[b][k] = 0.5f * (mx2[k] - mn2[k]); }
      }
      static int st6[1 << 16][2];
      int sp = 0; long lst = 0, iters = 0;
      st6[0][0] = 0; st6[0][1] = nent; sp = 1;
      while (sp > 0) {
        int take = sp < 32 ? sp : 32;
        int bf[32], be[32];
        for (int t = 0; t < take; t++) { --sp; bf[t] = st6[sp][0]; be[t] = st6[sp][1]; }
        iters++;
        for (int t = take - 1; t >= 0; t--) if (eskip[bf[t]] < be[t]) { st6[sp][0] = eskip[bf[t]]; st6[sp][1] = be[t]; sp++; }
        for (int t = take - 1; t >= 0; t--) {
          int j = bf[t];
          float dmin = 1e30f;
          for (int b = 0; b < 2; b++) {
            float d2 = 0;
            for (int k = 0; k < 3; k++) { float d = fabsf(ecen[j][k] - c2[b][k]) - h2[b][k]; if (d > 0) d2 += d * d; }
            if (d2 < dmin) dmin = d2;
          }
          if (ecen[j][3] < dmin) lst++;
          else { st6[sp][0] = j + 1; st6[sp][1] = eskip[j]; sp++; }
        }
      }
      lists6[ngroups] = lst; sum_l6 += lst; sum_it6 += iters;
    }
    lists[ngroups] = list;
    // (3) two-level: lane 0 first checks the first entry alone?  modelled as: windows in which
    // the first entry is accepted with skip beyond the window cost a "cheap" iteration
    ngroups++;
  }
  printf("G=%d W=%d theta=%.2f groups=%ld\n", G, W, theta, ngroups);
  printf("per-target accepted (reference criterion)  %.1f\n", sum_acc / ngroups);
  printf("per-lane scan: entries stepped per group    %.1f\n", sum_union / ngroups);
  printf("group scan: windows per group               %.1f (max %.0f)\n", sum_win / ngroups, max_win);
  printf("group scan: list length per group           %.1f (max %.0f)\n", sum_list / ngroups, max_list);
  printf("group scan: entries tested per group        %.1f\n", sum_tested / ngroups);
  printf("stack scan: iterations per group           %.1f, tested %.1f, max stack %.0f\n", sum_it4 / ngroups, sum_t4 / ngroups, max_sp4);
  printf("chain stack: iterations per group          %.1f, max stack %.0f\n", sum_it5 / ngroups, max_sp5);
  {
    double c = 0; long nab = 0; const long T = 2400;
    for (long a = 0; a < ngroups; a++) { if (lists[a] > T) { nab++; c += T * 8 + 35000 + its5[a] * 50.0 * T / lists[a]; } else c += lists[a] * 8 + its5[a] * 50; }
    printf("cost model (T=%ld): %.0f instr/group, %.2f%% aborted\n", T, c / ngroups, 100.0 * nab / ngroups);
    long big = 0; for (long a = 0; a < ngroups; a++) if (sps5[a] > 480) big++;
    printf("groups with chain stack > 480: %ld\n", big);
  }
  {
    for (long a = 1; a < ngroups; a++) { long v = lists6[a]; long b = a - 1; while (b >= 0 && lists6[b] > v) { lists6[b + 1] = lists6[b]; b--; } lists6[b + 1] = v; }
    long nab = 0; for (long a = 0; a < ngroups; a++) if (lists6[a] > 2400) nab++;
    printf("two-box: list mean %.1f iters %.1f p10 %ld p50 %ld p90 %ld p99 %ld max %ld  >2400: %.2f%%\n", sum_l6 / ngroups, sum_it6 / ngroups,
           lists6[ngroups / 10], lists6[ngroups / 2], lists6[ngroups * 9 / 10], lists6[ngroups * 99 / 100], lists6[ngroups - 1], 100.0 * nab / ngroups);
  }
  // percentiles of list length
  for (long a = 1; a < ngroups; a++) { long v = lists[a]; long b = a - 1; while (b >= 0 && lists[b] > v) { lists[b + 1] = lists[b]; b--; } lists[b + 1] = v; }
  printf("list length p10 %ld p50 %ld p90 %ld p99 %ld\n", lists[ngroups / 10], lists[ngroups / 2], lists[ngroups * 9 / 10], lists[ngroups * 99 / 100]);
  return 0;
}
