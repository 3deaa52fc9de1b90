/*
 * gh_oracle.c -- CPU restatement of GravHopper's force path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for gravhopper_b200.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.  The product path (gravhopper_b200/) never imports, links or executes it.
 *
 * Every function restates, in plain C with the same IEEE-754 double operations in the same
 * order, one function of the reference's C backend (/root/reference/gravhopper/_jbgrav.c,
 * v1.2.0) or one method of its Python step loop (/root/reference/gravhopper/gravhopper.py).
 * The restatement is pinned (tests/test_oracle.py, tests/golden/) against the reference's own
 * compiled C extension (oracle/_ref, built from the untouched sources by oracle/Makefile):
 * direct and tree results are BITWISE equal to the reference's on the golden inputs.
 *
 * Differences from the reference that do not change results:
 *   - no O(N^2) scratch arrays (the reference stores every pair before summing; the values
 *     summed, and their order, are identical -- see gho_direct below);
 *   - 64-bit indices (the reference's int arithmetic overflows for N > 26,754);
 *   - tree nodes come from one arena instead of one malloc each; OOM returns an error code
 *     instead of exit(209|435);
 *   - the per-target loops can run on several threads (OpenMP); each target's arithmetic is
 *     untouched, so results do not depend on the thread count.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no -ffast-math, ever).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GHO_OK 0
#define GHO_ENOMEM 1
#define GHO_EDEPTH 2

static void gho_set_threads(int nthreads)
{
#ifdef _OPENMP
	if (nthreads > 0) omp_set_num_threads(nthreads);
#else
	(void)nthreads;
#endif
}

int gho_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Direct summation on the particles themselves.
 * Follows directsummation_workhorse, _jbgrav.c:140-193.
 *
 * Reference pass 1 (:155-171) stores, for i<j, diff = x_i - x_j, dpos[i][j] = -diff,
 * dpos[j][i] = +diff, and w = 1.0 / s / sqrt(s) with s = (((0+d0^2)+d1^2)+d2^2) + eps^2 for
 * both (i,j) and (j,i).  -(x_i - x_j) == x_j - x_i exactly in IEEE arithmetic, and d^2 does
 * not depend on the sign, so dpos[i][j] == x_j - x_i and w[i][j] == w[j][i] bit for bit
 * whichever of i,j is smaller (only the sign of an exact zero can differ, which cannot
 * change a sum that starts from +0.0 ... except to keep it +0.0; see test_oracle).
 * Reference pass 2 (:174-185): acc starts at 0.0 and adds (mass[j]*dpos[i][j][k])*w[i][j]
 * for j = 0..N-1, j != i, strictly in that order.  No zero guard (:166).
 * ------------------------------------------------------------------------------------------ */
int gho_direct(const double *pos, const double *mass, int64_t np, double eps, double *acc,
               int nthreads)
{
	const double eps2 = eps * eps; /* :152 */
	gho_set_threads(nthreads);
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < np; i++) {
		double a0 = 0.0, a1 = 0.0, a2 = 0.0; /* :177 */
		const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
		for (int64_t j = 0; j < np; j++) {
			if (i == j) continue; /* :179 */
			/* value stored at dpos[i][j][k] (:160-161) */
			double d0, d1, d2;
			if (i < j) {
				d0 = -(xi - pos[3 * j]);
				d1 = -(yi - pos[3 * j + 1]);
				d2 = -(zi - pos[3 * j + 2]);
			} else {
				d0 = pos[3 * j] - xi;
				d1 = pos[3 * j + 1] - yi;
				d2 = pos[3 * j + 2] - zi;
			}
			double dpos2 = 0.0; /* :157 */
			dpos2 += d0 * d0;   /* :159,163 */
			dpos2 += d1 * d1;
			dpos2 += d2 * d2;
			const double s = dpos2 + eps2;           /* :165 */
			const double w = 1.0 / s / sqrt(s);      /* :166 */
			const double m = mass[j];
			a0 += m * d0 * w; /* :181-182, evaluated (m*d)*w */
			a1 += m * d1 * w;
			a2 += m * d2 * w;
		}
		acc[3 * i] = a0;
		acc[3 * i + 1] = a1;
		acc[3 * i + 2] = a2;
	}
	return GHO_OK;
}

/* ------------------------------------------------------------------------------------------
 * Direct summation at arbitrary target positions.
 * Follows directsummation_position_workhorse, _jbgrav.c:299-353: diff = target - source,
 * stored value -diff (:318), zero guard s == 0.0 -> w = 0 (:327-330), no j is skipped
 * (:343-346).
 * ------------------------------------------------------------------------------------------ */
int gho_direct_position(const double *pos, const double *mass, int64_t np, const double *fpos,
                        int64_t nf, double eps, double *acc, int nthreads)
{
	const double eps2 = eps * eps; /* :311 */
	gho_set_threads(nthreads);
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < nf; i++) {
		double a0 = 0.0, a1 = 0.0, a2 = 0.0; /* :342 */
		const double xi = fpos[3 * i], yi = fpos[3 * i + 1], zi = fpos[3 * i + 2];
		for (int64_t j = 0; j < np; j++) {
			const double d0 = -(xi - pos[3 * j]); /* :316,318 */
			const double d1 = -(yi - pos[3 * j + 1]);
			const double d2 = -(zi - pos[3 * j + 2]);
			double dpos2 = 0.0;
			dpos2 += d0 * d0;
			dpos2 += d1 * d1;
			dpos2 += d2 * d2;
			const double s = dpos2 + eps2; /* :321 */
			double w;
			if (s == 0.0) w = 0.0; /* :327-328 */
			else w = 1.0 / s / sqrt(s); /* :330 */
			const double m = mass[j];
			a0 += m * d0 * w; /* :344-345 */
			a1 += m * d1 * w;
			a2 += m * d2 * w;
		}
		acc[3 * i] = a0;
		acc[3 * i + 1] = a1;
		acc[3 * i + 2] = a2;
	}
	return GHO_OK;
}

/* ------------------------------------------------------------------------------------------
 * Barnes-Hut octree.  Node layout follows struct gravoct_node (_jbgrav.h:14-22); the fields
 * the reference never reads (boxmin/boxmax, COMvalid) are dropped, child pointers are arena
 * indices, and the leaf pointer is a particle index.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
	double center[3];
	double size, halfsize;
	double mass;
	double firstmoment[3];
	double COM[3];
	int64_t branches[8]; /* -1 = NULL */
	int64_t leaf;        /* particle index, -1 = NULL */
	int empty;
} gho_node;

typedef struct {
	gho_node *nodes;
	int64_t n, cap;
	const double *pos, *mass;
	int err;
	int64_t maxdepth;
} gho_tree;

/* gravoct_init, _jbgrav.c:360-384 */
static int64_t gho_node_new(gho_tree *t, const double *center, double size)
{
	if (t->n == t->cap) {
		int64_t ncap = t->cap * 2;
		gho_node *nn = (gho_node *)realloc(t->nodes, sizeof(gho_node) * (size_t)ncap);
		if (!nn) { t->err = GHO_ENOMEM; return -1; }
		t->nodes = nn;
		t->cap = ncap;
	}
	gho_node *r = &t->nodes[t->n];
	r->size = size;            /* :366 */
	r->halfsize = 0.5 * size;  /* :367 */
	for (int i = 0; i < 3; i++) {
		r->center[i] = center[i];
		r->firstmoment[i] = 0.0;
		r->COM[i] = 0.0;
	}
	for (int i = 0; i < 8; i++) r->branches[i] = -1;
	r->mass = 0.0;
	r->empty = 1;
	r->leaf = -1;
	return t->n++;
}

/* gravoct_calc_subnode + gravoct_calc_branchnum, _jbgrav.c:441-462: per axis +1 iff
 * p > centre (strict); branch bit k set iff axis k is +. */
static int gho_branch(const gho_node *nd, const double *p, int *subnode)
{
	int b = 0;
	for (int i = 0; i < 3; i++) {
		if (p[i] > nd->center[i]) { subnode[i] = 1; b += (1 << i); }
		else subnode[i] = -1;
	}
	return b;
}

/* gravoct_add_particle, _jbgrav.c:387-437.  Recursion depth is capped (the reference
 * recurses forever on coincident particles; we return GHO_EDEPTH instead of crashing). */
#define GHO_MAXDEPTH 200
static void gho_add(gho_tree *t, int64_t ni, int64_t p, int depth)
{
	if (t->err) return;
	if (depth > GHO_MAXDEPTH) { t->err = GHO_EDEPTH; return; }
	if (depth > t->maxdepth) t->maxdepth = depth;
	const double *pp = &t->pos[3 * p];
	const double pm = t->mass[p];
	int subnode[3];
	double subcenter[3];
	gho_node *nd = &t->nodes[ni];
	if (nd->empty) { /* :392-400 */
		nd->empty = 0;
		nd->leaf = p;
		nd->mass = pm;
		for (int i = 0; i < 3; i++) nd->firstmoment[i] = pm * pp[i];
	} else if (nd->leaf >= 0) { /* :401-416 */
		int64_t old = nd->leaf;
		int bnum = gho_branch(nd, &t->pos[3 * old], subnode);
		for (int i = 0; i < 3; i++) subcenter[i] = nd->center[i] + subnode[i] * 0.5 * nd->halfsize; /* :406 */
		int64_t c = gho_node_new(t, subcenter, nd->halfsize); /* :409 */
		if (c < 0) return;
		nd = &t->nodes[ni]; /* arena may have moved */
		nd->branches[bnum] = c;
		gho_add(t, c, old, depth + 1); /* :410 */
		nd = &t->nodes[ni];
		nd->leaf = -1;                 /* :411 */
		gho_add(t, ni, p, depth);      /* :413 */
	} else { /* :417-436 */
		int bnum = gho_branch(nd, pp, subnode);
		if (nd->branches[bnum] >= 0) {
			gho_add(t, nd->branches[bnum], p, depth + 1); /* :423 */
		} else {
			for (int i = 0; i < 3; i++) subcenter[i] = nd->center[i] + subnode[i] * 0.5 * nd->halfsize; /* :426-427 */
			int64_t c = gho_node_new(t, subcenter, nd->halfsize);
			if (c < 0) return;
			nd = &t->nodes[ni];
			nd->branches[bnum] = c;
			gho_add(t, c, p, depth + 1); /* :430 */
		}
		nd = &t->nodes[ni];
		nd->mass += pm; /* :432-435 */
		for (int i = 0; i < 3; i++) nd->firstmoment[i] += pm * pp[i];
	}
}

/* gravoct_finalize, _jbgrav.c:467-483, done eagerly for every node after the build (the
 * reference does it lazily on first acceptance; same arithmetic, same value). */
static void gho_finalize_all(gho_tree *t)
{
	for (int64_t k = 0; k < t->n; k++) {
		gho_node *nd = &t->nodes[k];
		if (nd->leaf >= 0) {
			for (int i = 0; i < 3; i++) nd->COM[i] = t->pos[3 * nd->leaf + i]; /* :473-475 */
		} else {
			for (int i = 0; i < 3; i++) nd->COM[i] = nd->firstmoment[i] / nd->mass; /* :477-479 */
		}
	}
}

/* gravoct_calc_accel, _jbgrav.c:487-541.  counters[0] += accepted nodes, [1] += visited. */
static void gho_accel(const gho_tree *t, int64_t ni, const double *pos, double eps, double theta,
                      double *force, int64_t *counters)
{
	const gho_node *nd = &t->nodes[ni];
	const double eps2 = eps * eps; /* :494 */
	double node_dist = 0.0;
	for (int i = 0; i < 3; i++)
		node_dist += (nd->center[i] - pos[i]) * (nd->center[i] - pos[i]); /* :496-499 */
	node_dist = sqrt(node_dist); /* :500 */
	if (counters) counters[1]++;
	if ((nd->leaf >= 0) || ((nd->size / node_dist) < theta)) { /* :502 */
		double d_pos[3], dpos2 = 0.0, invdpos3;
		for (int i = 0; i < 3; i++) {
			double diff = nd->COM[i] - pos[i]; /* :508 */
			d_pos[i] = diff;
			dpos2 += diff * diff;
		}
		double s = dpos2 + eps2; /* :513 */
		if (s == 0.0) invdpos3 = 0.0;          /* :517-518 */
		else invdpos3 = 1.0 / s / sqrt(s);     /* :520 */
		for (int i = 0; i < 3; i++) force[i] = d_pos[i] * nd->mass * invdpos3; /* :522-524 */
		if (counters) counters[0]++;
	} else {
		double branchforce[3];
		for (int i = 0; i < 3; i++) force[i] = 0.0; /* :527-529 */
		for (int j = 0; j < 8; j++) { /* :530-537 */
			if (nd->branches[j] >= 0) {
				gho_accel(t, nd->branches[j], pos, eps, theta, branchforce, counters);
				for (int i = 0; i < 3; i++) force[i] += branchforce[i];
			}
		}
	}
}

/* treeforce_workhorse, _jbgrav.c:737-806.  stats (nullable) receives
 * {nodes, max insert depth, accepted nodes summed over targets, visited nodes summed}. */
int gho_tree_force(const double *pos, const double *mass, int64_t np, const double *fpos,
                   int64_t nf, double eps, double theta, double *acc, int64_t *stats,
                   int nthreads)
{
	double min[3], max[3], boxsize, boxcenter[3];
	if (np < 1) return GHO_OK;
	for (int i = 0; i < 3; i++) { min[i] = pos[i]; max[i] = min[i]; } /* :747-750 */
	for (int64_t i = 1; i < np; i++) { /* :751-763 */
		for (int j = 0; j < 3; j++) {
			double q = pos[3 * i + j];
			if (q < min[j]) min[j] = q;
			if (q > max[j]) max[j] = q;
		}
	}
	boxsize = max[0] - min[0] + eps; /* :764 */
	for (int i = 1; i < 3; i++) {    /* :765-769: un-padded extent vs padded running value */
		if ((max[i] - min[i]) > boxsize) boxsize = max[i] - min[i] + eps;
	}
	for (int i = 0; i < 3; i++) boxcenter[i] = 0.5 * (min[i] + max[i]); /* :770-772 */

	gho_tree t;
	t.cap = 2 * np + 64;
	t.n = 0;
	t.pos = pos;
	t.mass = mass;
	t.err = 0;
	t.maxdepth = 0;
	t.nodes = (gho_node *)malloc(sizeof(gho_node) * (size_t)t.cap);
	if (!t.nodes) return GHO_ENOMEM;
	int64_t root = gho_node_new(&t, boxcenter, boxsize); /* :775 */
	for (int64_t i = 0; i < np && !t.err; i++) gho_add(&t, root, i, 0); /* :776-786 */
	if (t.err) { free(t.nodes); return t.err; }
	gho_finalize_all(&t);

	int64_t acc_cnt = 0, vis_cnt = 0;
	gho_set_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : acc_cnt, vis_cnt)
	for (int64_t i = 0; i < nf; i++) { /* :789-798 */
		double f[3];
		int64_t c[2] = {0, 0};
		gho_accel(&t, root, &fpos[3 * i], eps, theta, f, c);
		acc[3 * i] = f[0];
		acc[3 * i + 1] = f[1];
		acc[3 * i + 2] = f[2];
		acc_cnt += c[0];
		vis_cnt += c[1];
	}
	if (stats) { stats[0] = t.n; stats[1] = t.maxdepth; stats[2] = acc_cnt; stats[3] = vis_cnt; }
	free(t.nodes); /* :801 */
	return GHO_OK;
}

/* ------------------------------------------------------------------------------------------
 * CPU MODEL OF THE PRODUCT'S OPT-IN QUADRUPOLE EXTENSION.  NOT a reference
 * function: the reference evaluates monopoles only (_jbgrav.c:522-524).  Same octree, same
 * opening test (_jbgrav.c:502) and so the same accepted node set as gho_accel; every accepted
 * CELL additionally contributes its traceless quadrupole about its centre of mass,
 *   Q_ij = sum_k m_k (3 r_ki r_kj - r_k^2 delta_ij),  r_k = x_k - COM,
 *   a += -(Q e) y^5 + 5/2 (e.Q.e) y^7 e,   e = COM - x,  y = (|e|^2 + eps^2)^-1/2
 * (the gradient of -1/2 e.Q.e / r^5 with r^2 softened like the monopole).  Leaves have Q = 0.
 * The central second moments are built bottom-up with the parallel-axis rule, in double. */
static void gho_quad_build(const gho_tree *t, int64_t ni, double *C /* 6 per node: xx yy zz xy xz yz */)
{
	const gho_node *nd = &t->nodes[ni];
	double *c = &C[6 * ni];
	for (int k = 0; k < 6; k++) c[k] = 0.0;
	if (nd->leaf >= 0) return;
	for (int j = 0; j < 8; j++) {
		const int64_t ch = nd->branches[j];
		if (ch < 0) continue;
		gho_quad_build(t, ch, C);
		const gho_node *cn = &t->nodes[ch];
		const double dx = cn->COM[0] - nd->COM[0], dy = cn->COM[1] - nd->COM[1], dz = cn->COM[2] - nd->COM[2];
		const double *cc = &C[6 * ch];
		c[0] += cc[0] + cn->mass * dx * dx;
		c[1] += cc[1] + cn->mass * dy * dy;
		c[2] += cc[2] + cn->mass * dz * dz;
		c[3] += cc[3] + cn->mass * dx * dy;
		c[4] += cc[4] + cn->mass * dx * dz;
		c[5] += cc[5] + cn->mass * dy * dz;
	}
}

static void gho_accel_quad(const gho_tree *t, const double *C, int64_t ni, const double *pos, double eps,
                           double theta, double *force)
{
	const gho_node *nd = &t->nodes[ni];
	const double eps2 = eps * eps;
	double node_dist = 0.0;
	for (int i = 0; i < 3; i++) node_dist += (nd->center[i] - pos[i]) * (nd->center[i] - pos[i]);
	node_dist = sqrt(node_dist);
	if ((nd->leaf >= 0) || ((nd->size / node_dist) < theta)) {
		double e[3], s = eps2;
		for (int i = 0; i < 3; i++) { e[i] = nd->COM[i] - pos[i]; s += e[i] * e[i]; }
		const double y = (s == 0.0) ? 0.0 : 1.0 / sqrt(s);
		const double y3 = y * y * y;
		for (int i = 0; i < 3; i++) force[i] = e[i] * nd->mass * y3;
		if (nd->leaf < 0) {
			const double *c = &C[6 * ni];
			const double tr = c[0] + c[1] + c[2];
			const double qxx = 3.0 * c[0] - tr, qyy = 3.0 * c[1] - tr, qzz = 3.0 * c[2] - tr;
			const double qxy = 3.0 * c[3], qxz = 3.0 * c[4], qyz = 3.0 * c[5];
			const double qe[3] = {qxx * e[0] + qxy * e[1] + qxz * e[2], qxy * e[0] + qyy * e[1] + qyz * e[2],
			                      qxz * e[0] + qyz * e[1] + qzz * e[2]};
			const double eqe = e[0] * qe[0] + e[1] * qe[1] + e[2] * qe[2];
			const double y5 = y3 * y * y, y7 = y5 * y * y;
			for (int i = 0; i < 3; i++) force[i] += -qe[i] * y5 + 2.5 * eqe * y7 * e[i];
		}
	} else {
		double bf[3];
		for (int i = 0; i < 3; i++) force[i] = 0.0;
		for (int j = 0; j < 8; j++)
			if (nd->branches[j] >= 0) {
				gho_accel_quad(t, C, nd->branches[j], pos, eps, theta, bf);
				for (int i = 0; i < 3; i++) force[i] += bf[i];
			}
	}
}

int gho_tree_force_quad(const double *pos, const double *mass, int64_t np, const double *fpos, int64_t nf,
                        double eps, double theta, double *acc, int nthreads)
{
	double min[3], max[3], boxsize, boxcenter[3];
	if (np < 1) return GHO_OK;
	for (int i = 0; i < 3; i++) { min[i] = pos[i]; max[i] = min[i]; }
	for (int64_t i = 1; i < np; i++)
		for (int j = 0; j < 3; j++) {
			double q = pos[3 * i + j];
			if (q < min[j]) min[j] = q;
			if (q > max[j]) max[j] = q;
		}
	boxsize = max[0] - min[0] + eps;
	for (int i = 1; i < 3; i++)
		if ((max[i] - min[i]) > boxsize) boxsize = max[i] - min[i] + eps;
	for (int i = 0; i < 3; i++) boxcenter[i] = 0.5 * (min[i] + max[i]);
	gho_tree t;
	t.cap = 2 * np + 64; t.n = 0; t.pos = pos; t.mass = mass; t.err = 0; t.maxdepth = 0;
	t.nodes = (gho_node *)malloc(sizeof(gho_node) * (size_t)t.cap);
	if (!t.nodes) return GHO_ENOMEM;
	int64_t root = gho_node_new(&t, boxcenter, boxsize);
	for (int64_t i = 0; i < np && !t.err; i++) gho_add(&t, root, i, 0);
	if (t.err) { free(t.nodes); return t.err; }
	gho_finalize_all(&t);
	double *C = (double *)malloc(sizeof(double) * 6 * (size_t)t.n);
	if (!C) { free(t.nodes); return GHO_ENOMEM; }
	gho_quad_build(&t, root, C);
	gho_set_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t i = 0; i < nf; i++) {
		double f[3];
		gho_accel_quad(&t, C, root, &fpos[3 * i], eps, theta, f);
		acc[3 * i] = f[0]; acc[3 * i + 1] = f[1]; acc[3 * i + 2] = f[2];
	}
	free(C);
	free(t.nodes);
	return GHO_OK;
}

/* ------------------------------------------------------------------------------------------
 * CPU MODEL OF THE PRODUCT'S GROUP WALK (walk_group_kernel, gravhopper_b200/csrc/tree.cu).
 * This is NOT a reference function: the reference walks the tree once per target (gho_accel
 * above).  The product's default fp32 walk shares one traversal between the 32 depth-first
 * (Morton) consecutive targets of a warp and accepts a cell only if the reference's test
 * (_jbgrav.c:502) holds for every point of the targets' two bounding boxes.  The model restates
 * that criterion on the reference-shaped tree built above, so that tests can check the GPU kernel
 * against an independent implementation of ITS OWN criterion at full size, and can measure how
 * the criterion's error compares with the reference tree's.
 *
 * Decisions (cut of the group into two boxes, box extents with the kernel's padding, the
 * acceptance test) are made in single precision on (position - root centre), like the kernel;
 * forces are summed in double precision from the tree's double precision moments.  The traversal
 * reproduces the kernel's chain stack (pop up to 32 sibling chains per iteration, push the rest of
 * each chain, then the children of opened cells) so that the two give-up conditions -- list longer
 * than list_limit when a chunk is evaluated, more than stack_limit chains -- trigger for the same
 * groups; such groups use the per-target walk gho_accel, as the kernel does.
 *   order_out[np]  (nullable) depth-first leaf order (sorted position -> particle)
 *   list_out[np]   (nullable) per particle: list length of its group, -1 if the group gave up
 * ------------------------------------------------------------------------------------------ */
typedef struct { float cx, cy, cz, hx, hy, hz; } gho_box;

static gho_box gho_box_of(const float *x, const float *y, const float *z, int a, int b)
{
	gho_box bx;
	float lx = INFINITY, ux = -INFINITY, ly = INFINITY, uy = -INFINITY, lz = INFINITY, uz = -INFINITY;
	for (int l = a; l < b; l++) {
		lx = fminf(lx, x[l]); ux = fmaxf(ux, x[l]);
		ly = fminf(ly, y[l]); uy = fmaxf(uy, y[l]);
		lz = fminf(lz, z[l]); uz = fmaxf(uz, z[l]);
	}
	const float pad = 1.0f + 1e-6f;
	bx.cx = 0.5f * (lx + ux); bx.cy = 0.5f * (ly + uy); bx.cz = 0.5f * (lz + uz);
	bx.hx = (0.5f * (ux - lx)) * pad + 1e-6f * fabsf(bx.cx);
	bx.hy = (0.5f * (uy - ly)) * pad + 1e-6f * fabsf(bx.cy);
	bx.hz = (0.5f * (uz - lz)) * pad + 1e-6f * fabsf(bx.cz);
	return bx;
}

static float gho_box_dist2(const gho_box *b, float cx, float cy, float cz)
{
	const float dx = fmaxf(fabsf(cx - b->cx) - b->hx, 0.f);
	const float dy = fmaxf(fabsf(cy - b->cy) - b->hy, 0.f);
	const float dz = fmaxf(fabsf(cz - b->cz) - b->hz, 0.f);
	return dx * dx + dy * dy + dz * dz;
}

typedef struct { int64_t parent; int slot; } gho_chain; /* siblings = children of parent at slots >= slot */

int gho_tree_force_group2(const double *pos, const double *mass, int64_t np, double eps, double theta,
                          int list_limit, int stack_limit, double *acc, int64_t *order_out,
                          int32_t *list_out, int64_t *stats, int nthreads, const int64_t *order_in,
                          int group_size, int two_boxes, double *abs_out, double hybrid_kappa);

int gho_tree_force_group(const double *pos, const double *mass, int64_t np, double eps, double theta,
                         int list_limit, int stack_limit, double *acc, int64_t *order_out,
                         int32_t *list_out, int64_t *stats, int nthreads)
{
	return gho_tree_force_group2(pos, mass, np, eps, theta, list_limit, stack_limit, acc, order_out,
	                             list_out, stats, nthreads, NULL, 32, 1, NULL, 0.0);
}

/* Design-study form of the model: the targets can be grouped along any given order (order_in, e.g.
 * a Hilbert curve instead of the depth-first/Morton order), in groups of group_size <= 64, with one
 * or two bounding boxes.  gho_tree_force_group is the kernel's configuration (NULL, 32, 1).
 * abs_out[np] (nullable) receives S = 16 x sum of m_e / (|d_e|^2 + eps^2) over the entries at
 * positions 0 and 1 of every 32-entry chunk of the target's list: the kernel's 1/16 sample of the
 * summed magnitude of the contributions, against which |acc| measures how strongly they cancel.
 * hybrid_kappa > 0: the kernel's hybrid rule -- targets with |acc| < kappa * S are re-evaluated
 * with the per-target walk (gho_accel); stats[6] counts them. */
int gho_tree_force_group2(const double *pos, const double *mass, int64_t np, double eps, double theta,
                          int list_limit, int stack_limit, double *acc, int64_t *order_out,
                          int32_t *list_out, int64_t *stats, int nthreads, const int64_t *order_in,
                          int group_size, int two_boxes, double *abs_out, double hybrid_kappa)
{
	double min[3], max[3], boxsize, boxcenter[3];
	if (np < 1) return GHO_OK;
	for (int i = 0; i < 3; i++) { min[i] = pos[i]; max[i] = min[i]; }
	for (int64_t i = 1; i < np; i++)
		for (int j = 0; j < 3; j++) {
			double q = pos[3 * i + j];
			if (q < min[j]) min[j] = q;
			if (q > max[j]) max[j] = q;
		}
	boxsize = max[0] - min[0] + eps;
	for (int i = 1; i < 3; i++)
		if ((max[i] - min[i]) > boxsize) boxsize = max[i] - min[i] + eps;
	for (int i = 0; i < 3; i++) boxcenter[i] = 0.5 * (min[i] + max[i]);

	gho_tree t;
	t.cap = 2 * np + 64; t.n = 0; t.pos = pos; t.mass = mass; t.err = 0; t.maxdepth = 0;
	t.nodes = (gho_node *)malloc(sizeof(gho_node) * (size_t)t.cap);
	if (!t.nodes) return GHO_ENOMEM;
	int64_t root = gho_node_new(&t, boxcenter, boxsize);
	for (int64_t i = 0; i < np && !t.err; i++) gho_add(&t, root, i, 0);
	if (t.err) { free(t.nodes); return t.err; }
	gho_finalize_all(&t);

	/* depth-first leaf order (children in branch order) and the level of every node */
	int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)np);
	signed char *level = (signed char *)malloc((size_t)t.n);
	int64_t *stk = (int64_t *)malloc(sizeof(int64_t) * (size_t)(8 * (GHO_MAXDEPTH + 2)));
	if (!order || !level || !stk) { free(order); free(level); free(stk); free(t.nodes); return GHO_ENOMEM; }
	{
		int64_t sp = 0, k = 0;
		stk[sp++] = root;
		level[root] = 0;
		while (sp > 0) {
			const int64_t ni = stk[--sp];
			const gho_node *nd = &t.nodes[ni];
			if (nd->leaf >= 0) { order[k++] = nd->leaf; continue; }
			for (int j = 7; j >= 0; j--)
				if (nd->branches[j] >= 0) {
					level[nd->branches[j]] = (signed char)(level[ni] + 1);
					stk[sp++] = nd->branches[j];
				}
		}
	}
	free(stk);
	if (order_out) memcpy(order_out, order, sizeof(int64_t) * (size_t)np);
	if (order_in) memcpy(order, order_in, sizeof(int64_t) * (size_t)np);
	if (group_size < 1 || group_size > 64) group_size = 32;
	const int GS = group_size;

	const double inv_theta2 = 1.0 / (theta * theta);
	const float s2root = (float)(boxsize * boxsize * inv_theta2);
	const double eps2 = eps * eps;
	const int64_t ngroups = (np + GS - 1) / GS;
	int64_t n_list = 0, n_tested = 0, n_iter = 0, n_fallback = 0, n_redo = 0;
	int oom = 0;
	gho_set_threads(nthreads);
#pragma omp parallel reduction(+ : n_list, n_tested, n_iter, n_fallback, n_redo)
	{
		int64_t cap = 4096, *lst = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
		gho_chain *stack = (gho_chain *)malloc(sizeof(gho_chain) * (size_t)(stack_limit + 64));
		if (!lst || !stack) oom = 1;
#pragma omp for schedule(dynamic, 16)
		for (int64_t g = 0; g < ngroups; g++) {
			if (oom) continue;
			const int64_t p0 = (int64_t)GS * g;
			const int nv = (int)((np - p0) < GS ? (np - p0) : GS);
			float x[64], y[64], z[64];
			for (int l = 0; l < nv; l++) {
				const double *q = &pos[3 * order[p0 + l]];
				x[l] = (float)(q[0] - boxcenter[0]);
				y[l] = (float)(q[1] - boxcenter[1]);
				z[l] = (float)(q[2] - boxcenter[2]);
			}
			int cut = 0;
			float gmax = -1.f;
			for (int l = 0; l + 1 < nv; l++) {
				const float gap = (x[l + 1] - x[l]) * (x[l + 1] - x[l]) + (y[l + 1] - y[l]) * (y[l + 1] - y[l]) +
				                  (z[l + 1] - z[l]) * (z[l + 1] - z[l]);
				if (gap > gmax) { gmax = gap; cut = l; }
			}
			if (!two_boxes) cut = nv - 1;
			const gho_box A = gho_box_of(x, y, z, 0, cut + 1);
			const gho_box B = (cut + 1 < nv) ? gho_box_of(x, y, z, cut + 1, nv) : A;

			int sp = 0, fallback = 0;
			int64_t head = 0, tail = 0;
			stack[sp].parent = -1; stack[sp].slot = 0; sp++;
			while (sp > 0 && !fallback) {
				n_iter++;
				const int take = sp < 32 ? sp : 32;
				gho_chain item[32];
				for (int l = 0; l < take; l++) item[l] = stack[sp - 1 - l];
				sp -= take;
				if (head - tail >= 32) {
					tail += 32;
					if (head > list_limit) { fallback = 1; break; }
				}
				gho_chain rem[32], opn[32];
				int nrem = 0, nopn = 0;
				int64_t accn[32];
				int nacc = 0;
				for (int l = 0; l < take; l++) {
					int64_t first;
					int k = item[l].slot, more = 0;
					if (item[l].parent < 0) {
						first = root;
					} else {
						const gho_node *pn = &t.nodes[item[l].parent];
						while (pn->branches[k] < 0) k++;
						first = pn->branches[k];
						for (int j = k + 1; j < 8; j++)
							if (pn->branches[j] >= 0) { more = 1; break; }
					}
					const gho_node *nd = &t.nodes[first];
					n_tested += nv;
					int accept;
					if (nd->leaf >= 0) {
						accept = 1;
					} else {
						const float cx = (float)(nd->center[0] - boxcenter[0]);
						const float cy = (float)(nd->center[1] - boxcenter[1]);
						const float cz = (float)(nd->center[2] - boxcenter[2]);
						const float s2 = s2root * ldexpf(1.0f, -2 * level[first]);
						const float d2 = fminf(gho_box_dist2(&A, cx, cy, cz), gho_box_dist2(&B, cx, cy, cz));
						accept = (s2 < d2);
					}
					if (more) { rem[nrem].parent = item[l].parent; rem[nrem].slot = k + 1; nrem++; }
					if (accept) accn[nacc++] = first;
					else { opn[nopn].parent = first; opn[nopn].slot = 0; nopn++; }
				}
				if (sp + nrem + nopn > stack_limit) { fallback = 1; break; }
				/* lane 0 held the top of the stack and its pushes end up on top again */
				for (int l = nrem - 1; l >= 0; l--) stack[sp++] = rem[l];
				for (int l = nopn - 1; l >= 0; l--) stack[sp++] = opn[l];
				if (head + nacc > cap) {
					cap *= 2;
					int64_t *nl = (int64_t *)realloc(lst, sizeof(int64_t) * (size_t)cap);
					if (!nl) { oom = 1; fallback = 1; break; }
					lst = nl;
				}
				for (int l = 0; l < nacc; l++) lst[head++] = accn[l];
			}
			for (int l = 0; l < nv; l++) {
				const int64_t pi = order[p0 + l];
				const double *q = &pos[3 * pi];
				double f[3] = {0.0, 0.0, 0.0}, fabs_sum = 0.0;
				if (fallback) {
					gho_accel(&t, root, q, eps, theta, f, NULL);
				} else {
					for (int64_t e = 0; e < head; e++) {
						const gho_node *nd = &t.nodes[lst[e]];
						const double dx = nd->COM[0] - q[0], dy = nd->COM[1] - q[1], dz = nd->COM[2] - q[2];
						const double s = dx * dx + dy * dy + dz * dz + eps2;
						const double w = (s == 0.0) ? 0.0 : nd->mass / s / sqrt(s);
						f[0] += dx * w; f[1] += dy * w; f[2] += dz * w;
						if ((e & 31) < 2) fabs_sum += nd->mass / s;
					}
				}
				fabs_sum *= 16.0;
				if (abs_out) abs_out[pi] = fabs_sum;
				if (!fallback && hybrid_kappa > 0.0 &&
				    f[0] * f[0] + f[1] * f[1] + f[2] * f[2] < hybrid_kappa * hybrid_kappa * fabs_sum * fabs_sum) {
					gho_accel(&t, root, q, eps, theta, f, NULL);
					n_redo++;
				}
				acc[3 * pi] = f[0]; acc[3 * pi + 1] = f[1]; acc[3 * pi + 2] = f[2];
				if (list_out) list_out[pi] = fallback ? -1 : (int32_t)head;
			}
			if (fallback) n_fallback++;
			else n_list += head * nv;
		}
		free(lst);
		free(stack);
	}
	if (stats) { stats[0] = t.n; stats[1] = n_list; stats[2] = n_tested; stats[3] = n_iter; stats[4] = n_fallback; stats[5] = ngroups; stats[6] = n_redo; }
	free(order); free(level); free(t.nodes);
	return oom ? GHO_ENOMEM : GHO_OK;
}

/* ------------------------------------------------------------------------------------------
 * One drift-kick-drift leapfrog step in the reference's internal units (kpc, km/s, Msun, Myr).
 * Follows Simulation.perform_timestep, gravhopper.py:405-416, with the unit handling of
 * jbgrav.py:38-48 made explicit:
 *   x_half = x + ((0.5*v)*dt) * KPC_PER_KMS_MYR          (:409; astropy scales the product)
 *   a      = C_ACC * force(x_half)  [+ ext]              (:413; jbgrav.py:48)
 *   v_new  = v + a*dt                                    (:414)
 *   x_new  = x_half + ((0.5*v_new)*dt) * KPC_PER_KMS_MYR (:416)
 * algorithm: 0 = direct (gho_direct), 1 = tree (gho_tree_force on the particles themselves,
 * theta as given; the reference always uses 0.7, gravhopper.py:446 + jbgrav.py:52).
 * ext (nullable): extra acceleration (km/s/Myr) evaluated by the caller at x_half.
 * xhalf_out (nullable) receives x_half.
 * ------------------------------------------------------------------------------------------ */
#define GHO_KPC_PER_KMS_MYR 1.022712165045695e-3
#define GHO_C_ACC 4.398600412921223e-09

void gho_half_drift(const double *x, const double *v, int64_t np, double dt, double *xhalf)
{
	for (int64_t q = 0; q < 3 * np; q++)
		xhalf[q] = x[q] + ((0.5 * v[q]) * dt) * GHO_KPC_PER_KMS_MYR;
}

int gho_leapfrog_step(double *x, double *v, const double *mass, int64_t np, double dt, double eps,
                      double theta, int algorithm, const double *ext, double *xhalf_out,
                      int nthreads)
{
	double *xh = (double *)malloc(sizeof(double) * 3 * (size_t)np);
	double *a = (double *)malloc(sizeof(double) * 3 * (size_t)np);
	if (!xh || !a) { free(xh); free(a); return GHO_ENOMEM; }
	int rc = GHO_OK;
	gho_half_drift(x, v, np, dt, xh);
	if (np > 1) { /* gravhopper.py:442 */
		if (algorithm == 0) rc = gho_direct(xh, mass, np, eps, a, nthreads);
		else rc = gho_tree_force(xh, mass, np, xh, np, eps, theta, a, NULL, nthreads);
	} else {
		for (int64_t q = 0; q < 3 * np; q++) a[q] = 0.0; /* gravhopper.py:449-450 */
	}
	if (rc == GHO_OK) {
		for (int64_t q = 0; q < 3 * np; q++) {
			double acc = a[q] * GHO_C_ACC;  /* jbgrav.py:48 */
			if (ext) acc = acc + ext[q];    /* gravhopper.py:457 */
			double vn = v[q] + acc * dt;    /* :414 */
			v[q] = vn;
			x[q] = xh[q] + ((0.5 * vn) * dt) * GHO_KPC_PER_KMS_MYR; /* :416 */
		}
		if (xhalf_out) memcpy(xhalf_out, xh, sizeof(double) * 3 * (size_t)np);
	}
	free(xh);
	free(a);
	return rc;
}

/* ------------------------------------------------------------------------------------------
 * Energy diagnostic.  The reference has no energy routine; this is the energy consistent with
 * its force law (_jbgrav.c:165-166): KE = 1/2 sum m v^2, PE = -G sum_{i<j} m_i m_j /
 * sqrt(r_ij^2 + eps^2), G = 4.30091727003628e-06 kpc (km/s)^2 / Msun.  out = {KE, PE}.
 * ------------------------------------------------------------------------------------------ */
#define GHO_G 4.30091727003628e-06
void gho_energy(const double *x, const double *v, const double *mass, int64_t np, double eps,
                double *out, int nthreads)
{
	double ke = 0.0, pe = 0.0;
	const double eps2 = eps * eps;
	gho_set_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ke, pe)
	for (int64_t i = 0; i < np; i++) {
		ke += 0.5 * mass[i] * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
		double p = 0.0;
		for (int64_t j = i + 1; j < np; j++) {
			double d0 = x[3 * j] - x[3 * i], d1 = x[3 * j + 1] - x[3 * i + 1], d2 = x[3 * j + 2] - x[3 * i + 2];
			p += mass[j] / sqrt(d0 * d0 + d1 * d1 + d2 * d2 + eps2);
		}
		pe -= GHO_G * mass[i] * p;
	}
	out[0] = ke;
	out[1] = pe;
}
