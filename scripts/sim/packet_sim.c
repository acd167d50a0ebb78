// Design-space simulator (CPU, not product): counts warp-level work of a 32-query packet traversal of the igl tree
// versus 32 independent igl-order traversals.  gcc -O2 -shared -fPIC -o /tmp/packet_sim.so packet_sim.c
#include <math.h>
#include <stdint.h>
#include <string.h>
typedef struct { double x, y, z; } V3;
static double dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 sub(V3 a, V3 b) { V3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static V3 cpt(V3 p, V3 a, V3 b, V3 c) {
	V3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
	double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
	if (d1 <= 0 && d2 <= 0) return a;
	V3 bp = sub(p, b); double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
	if (d3 >= 0 && d4 <= d3) return b;
	double vc = d1 * d4 - d3 * d2;
	if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); V3 r = {a.x + v * ab.x, a.y + v * ab.y, a.z + v * ab.z}; return r; }
	V3 cp = sub(p, c); double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
	if (d6 >= 0 && d5 <= d6) return c;
	double vb = d5 * d2 - d1 * d6;
	if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); V3 r = {a.x + w * ac.x, a.y + w * ac.y, a.z + w * ac.z}; return r; }
	double va = d3 * d6 - d5 * d4;
	if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); V3 r = {b.x + w * (c.x - b.x), b.y + w * (c.y - b.y), b.z + w * (c.z - b.z)}; return r; }
	double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
	V3 r = {a.x + ab.x * v + ac.x * w, a.y + ab.y * v + ac.y * w, a.z + ab.z * v + ac.z * w}; return r;
}
static double boxd(const double *b, V3 p) {
	double d = 0, a;
	if (b[0] > p.x) { a = b[0] - p.x; d += a * a; } else if (p.x > b[3]) { a = p.x - b[3]; d += a * a; }
	if (b[1] > p.y) { a = b[1] - p.y; d += a * a; } else if (p.y > b[4]) { a = p.y - b[4]; d += a * a; }
	if (b[2] > p.z) { a = b[2] - p.z; d += a * a; } else if (p.z > b[5]) { a = p.z - b[5]; d += a * a; }
	return d;
}
// tree: pre-order nodes, box[6n], prim[n] (>=0 for leaves), lr[2n]
// out[0]=sum single node visits, out[1]=sum single leaf tests, out[2]=packet node visits (warp level), out[3]=packet leaf steps,
// out[4]=sum of lanes active in packet leaf steps, out[5]=sum lanes wanting at packet node visits, out[6]=max-lane single nodes summed per warp
void simulate(const double *box, const int32_t *prim, const int32_t *lr, const double *tri, const double *P, int64_t np, int W, double *out) {
	memset(out, 0, 8 * sizeof(double));
	for (int64_t g = 0; g + W <= np; g += W) {
		// singles
		double mx = 0;
		for (int l = 0; l < W; ++l) {
			V3 p = {P[3 * (g + l)], P[3 * (g + l) + 1], P[3 * (g + l) + 2]};
			double best = INFINITY; int st[128]; double sd[128]; int sp = 0; st[sp] = 0; sd[sp++] = 0; double nv = 0;
			while (sp) {
				--sp; if (!(sd[sp] < best)) continue; int n = st[sp];
				if (prim[n] >= 0) { const double *t = tri + 9 * (int64_t)prim[n]; V3 a = {t[0], t[1], t[2]}, b = {t[3], t[4], t[5]}, c = {t[6], t[7], t[8]};
					V3 q = cpt(p, a, b, c); V3 d = sub(p, q); double dd = dot3(d, d); if (dd < best) best = dd; out[1] += 1; continue; }
				nv += 1;
				int L = lr[2 * n], R = lr[2 * n + 1]; double dl = boxd(box + 6 * L, p), dr = boxd(box + 6 * R, p);
				if (dl < dr || dl == 0) { st[sp] = R; sd[sp++] = dr; st[sp] = L; sd[sp++] = dl; } else { st[sp] = L; sd[sp++] = dl; st[sp] = R; sd[sp++] = dr; }
			}
			out[0] += nv; if (nv > mx) mx = nv;
		}
		out[6] += mx;
		// packet
		V3 p[64]; double best[64];
		for (int l = 0; l < W; ++l) { p[l].x = P[3 * (g + l)]; p[l].y = P[3 * (g + l) + 1]; p[l].z = P[3 * (g + l) + 2]; best[l] = INFINITY; }
		int st[256]; int sp = 0; st[sp++] = 0;
		while (sp) {
			int n = st[--sp];
			if (prim[n] >= 0) {
				int act = 0;
				for (int l = 0; l < W; ++l) if (boxd(box + 6 * n, p[l]) <= best[l]) {
					++act; const double *t = tri + 9 * (int64_t)prim[n]; V3 a = {t[0], t[1], t[2]}, b = {t[3], t[4], t[5]}, c = {t[6], t[7], t[8]};
					V3 q = cpt(p[l], a, b, c); V3 d = sub(p[l], q); double dd = dot3(d, d); if (dd < best[l]) best[l] = dd; }
				if (act) { out[3] += 1; out[4] += act; }
				continue;
			}
			// re-test on pop: any lane still wants this node?
			int want = 0; for (int l = 0; l < W; ++l) if (boxd(box + 6 * n, p[l]) <= best[l]) ++want;
			if (!want) continue;
			out[2] += 1; out[5] += want;
			int L = lr[2 * n], R = lr[2 * n + 1]; int wl = 0, wr = 0, nearl = 0;
			for (int l = 0; l < W; ++l) { double dl = boxd(box + 6 * L, p[l]), dr = boxd(box + 6 * R, p[l]); if (dl <= best[l]) ++wl; if (dr <= best[l]) ++wr; if (dl < dr || dl == 0) ++nearl; }
			if (2 * nearl >= W) { if (wr) st[sp++] = R; if (wl) st[sp++] = L; } else { if (wl) st[sp++] = L; if (wr) st[sp++] = R; }
		}
	}
}
