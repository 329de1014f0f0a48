/* rtb_render_cli.c — a C99 consumer of librtb200's C ABI (include/rtb200.h), no Python in between.
 *
 * Renders one frame the way the reference's RenderJob::run delivers it (src/server.rs:157-199): the streaming job yields
 * the 60-pixel records while the frame renders, this program paints them into a frame and writes a binary PPM — what
 * the reference's removed `--image <png> --spp <n> --scene <name>` CLI did (render_examples.sh:8).
 *
 *   cc -std=c99 -Iinclude examples/rtb_render_cli.c -Lraytracer-server_b200 -lrtb200 -Wl,-rpath,$PWD/raytracer-server_b200 -o rtb_render_cli
 *   ./rtb_render_cli <scene.toml> <assets dir> <width> <height> <spp> <out.ppm> [octree] [seed]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtb200.h"

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s <scene.toml> <assets dir> <width> <height> <spp> <out.ppm> [octree] [seed]\n", argv[0]);
        return 2;
    }
    rtb_scene* scene = NULL;
    int rc = rtb_scene_load_toml(argv[1], argv[2], 0, &scene);
    if (rc != RTB_OK) {   /* LoadTomlError::{Io, Parse, MeshLoad} and the reference's panics, as codes (src/scene.rs:350-355) */
        fprintf(stderr, "failed to load scene (%d): %s\n", rc, rtb_last_error());
        return 1;
    }
    rtb_params p;
    memset(&p, 0, sizeof(p));
    p.width = atoi(argv[3]);
    p.height = atoi(argv[4]);
    p.spp = atoi(argv[5]);
    p.world = 1;
    p.accel = argc > 7 && strcmp(argv[7], "octree") == 0 ? RTB_ACCEL_OCTREE_REFERENCE : RTB_ACCEL_LBVH;
    p.seed = argc > 8 ? strtoull(argv[8], NULL, 10) : 0;
    rtb_job* job = NULL;
    rc = rtb_job_begin(scene, &p, 1, &job);
    if (rc != RTB_OK) {
        fprintf(stderr, "rtb_job_begin (%d): %s\n", rc, rtb_last_error());
        rtb_scene_destroy(scene);
        return 1;
    }
    uint8_t* frame = (uint8_t*)calloc((size_t)p.width * p.height, 3);
    uint8_t buf[256 * (6 + 3 * 60)];
    long records = 0, pixels = 0;
    for (;;) {
        int64_t n = 0;
        rc = rtb_job_next_messages(job, buf, (int64_t)sizeof(buf), 256, &n);
        if (rc <= 0) break;   /* 0 = frame complete; negative = error / stopped */
        for (int64_t off = 0; off < n;) {   /* each record is one wire message: [0, n, x u16le, y u16le, n x rgb] (src/server.rs:173-190) */
            const int cnt = buf[off + 1], x = buf[off + 2] | (buf[off + 3] << 8), y = buf[off + 4] | (buf[off + 5] << 8);
            memcpy(frame + ((size_t)y * p.width + x) * 3, buf + off + 6, (size_t)cnt * 3);
            off += 6 + 3 * cnt;
            pixels += cnt;
            ++records;
        }
    }
    if (rc < 0) fprintf(stderr, "render failed (%d): %s\n", rc, rtb_last_error());
    rtb_stats st;
    memset(&st, 0, sizeof(st));
    rtb_job_stats(job, &st);
    const int stopped = rtb_job_end(job);
    FILE* f = fopen(argv[6], "wb");
    if (f) {
        fprintf(f, "P6\n%d %d\n255\n", p.width, p.height);
        fwrite(frame, 3, (size_t)p.width * p.height, f);
        fclose(f);
    }
    printf("%ld records, %ld pixels, %llu samples, %llu rays, first record after %.2f ms, frame after %.2f ms%s\n", records, pixels,
           (unsigned long long)st.samples, (unsigned long long)(st.rays_primary + st.rays_extension + st.rays_shadow), st.first_record_ms,
           st.wall_ms, stopped == RTB_ECANCELLED ? " (stopped early)" : "");
    free(frame);
    rtb_scene_destroy(scene);
    return rc < 0 || !f ? 1 : 0;
}
