// Prints the launch plans of the peak kernels for a few shapes (host only): nvcc -o /tmp/print_plan tools/print_plan.cu && /tmp/print_plan
#include <cstdio>
#include <cstring>
#include "../object_keypoints_b200/csrc/okp_peaks_stream.cuh"
int main() {
    const int shapes[][3] = {{180, 320, 4096 * 3}, {64, 64, 32768 * 3}, {64, 64, 256 * 3}, {128, 128, 3000}, {360, 480, 300}};
    for (auto& sh : shapes)
        for (int esize = 4; esize >= 2; esize -= 2)
            for (int fused = 0; fused < 2; ++fused) {
                OkpStreamPlan sp;
                const int frame_bytes = fused ? (int)okp_group_smem_bytes(3, 32, 16) : 0;
                if (!okp_stream_plan(sh[2], 3, sh[0], sh[1], 32, esize, frame_bytes, 0, &sp)) { printf("%dx%d esize %d: no plan\n", sh[0], sh[1], esize); continue; }
                printf("%3dx%3d esize %d fused %d: M %2d strips %2d compute threads %3d EW %d total threads %3d smem %6d NS %d nb %d groups %d F %d\n",
                       sh[0], sh[1], esize, fused, sp.s.M, sp.s.strips, sp.s.threads, sp.EW, sp.threads, sp.smem_bytes, sp.s.NS, sp.s.nb, sp.groups, sp.F);
            }
    return 0;
}
