/*
 * oracle/voxelize.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the hard voxelizer that the reference calls at
 *   pcdet/datasets/processor/data_processor.py L54
 *       self._voxel_generator.point_to_voxel(tv.from_numpy(points))      (spconv 2.x)
 *   pcdet/datasets/processor/data_processor.py L46  generate(points)     (spconv 1.x)
 * The loop itself lives in the un-vendored, un-pinned third-party package
 * `spconv` (spconv.utils.Point2VoxelCPU3d); it is restated here from its
 * published behaviour.  PARITY UNPINNED by the
 * reference (it ships no tests); pinned by tests/test_oracle_voxelize.py
 * against a literal pure-Python loop and by the golden vectors produced
 * through the reference's own data_processor.py (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* One frame.  points: (n, f) float32 row-major, columns 0..2 are x,y,z.
 * vsize_xyz[3], range_xyz[6] (xmin,ymin,zmin,xmax,ymax,zmax), all float32
 * as the spconv generator stores them.  grid_xyz[3] = round((hi-lo)/vsize),
 * computed by the caller exactly as data_processor.py L117-118 does.
 * table: caller-provided int32 scratch of grid_z*grid_y*grid_x entries, all -1
 * on entry; restored to -1 on exit (so repeated calls do not pay a memset of
 * the 85M-entry table -- the published generator does the same reset).
 * Outputs (capacity max_voxels): voxels (max_voxels, max_points, f) must be
 * zero on entry; coords (max_voxels, 3) in z,y,x; num (max_voxels) zero.
 * legacy_break != 0 reproduces spconv 1.x (break when the voxel budget is hit);
 * 0 is the spconv 2.x behaviour (skip that point, keep scanning).
 * Returns the number of voxels produced. */
int oracle_points_to_voxels(const float *points, int64_t n, int f,
                            const float *vsize_xyz, const float *range_xyz,
                            const int *grid_xyz, int max_points, int max_voxels,
                            int legacy_break, int32_t *table,
                            float *voxels, int32_t *coords, int32_t *num)
{
    const int gx = grid_xyz[0], gy = grid_xyz[1], gz = grid_xyz[2];
    int nvox = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = points + i * f;
        int c[3]; /* c[0]=z, c[1]=y, c[2]=x */
        int ok = 1;
        for (int j = 0; j < 3; ++j) { /* j over x,y,z as the generator does */
            volatile float d = p[j] - range_xyz[j];      /* fp32 subtract  */
            volatile float q = d / vsize_xyz[j];         /* fp32 true divide */
            float fl = floorf(q);
            int g = (j == 0) ? gx : (j == 1 ? gy : gz);
            /* compare in float first so that huge / NaN values do not hit UB in the cast */
            if (!(fl >= 0.0f) || !(fl < (float)g)) { ok = 0; break; }
            c[2 - j] = (int)fl;
        }
        if (!ok) continue;
        int64_t cell = ((int64_t)c[0] * gy + c[1]) * gx + c[2];
        int v = table[cell];
        if (v == -1) {
            if (nvox >= max_voxels) {
                if (legacy_break) break;
                continue;
            }
            v = nvox++;
            table[cell] = v;
            coords[v * 3 + 0] = c[0];
            coords[v * 3 + 1] = c[1];
            coords[v * 3 + 2] = c[2];
        }
        if (num[v] < max_points) {
            memcpy(voxels + ((int64_t)v * max_points + num[v]) * f, p, sizeof(float) * f);
            num[v] += 1;
        }
    }
    for (int v = 0; v < nvox; ++v) {
        int64_t cell = ((int64_t)coords[v * 3] * gy + coords[v * 3 + 1]) * gx + coords[v * 3 + 2];
        table[cell] = -1;
    }
    return nvox;
}

/* MeanVFE on the padded tensor: pcdet/models/backbones_3d/vfe/mean_vfe.py L25-29.
 * Sum over ALL max_points slots (padding is zero) divided by max(num,1). */
void oracle_mean_vfe(const float *voxels, const int32_t *num, int64_t nvox,
                     int max_points, int f, float *out)
{
    for (int64_t v = 0; v < nvox; ++v) {
        float nrm = num[v] < 1 ? 1.0f : (float)num[v];
        for (int c = 0; c < f; ++c) {
            float s = 0.0f;
            for (int k = 0; k < max_points; ++k)
                s += voxels[(v * max_points + k) * f + c];
            out[v * f + c] = s / nrm;
        }
    }
}
