// Development aid (not product, not a fallback): runs the __host__ __device__ arithmetic of
// csrc/okp_stereo.cuh on the CPU of the GPU-less build container.
//   nvcc -O2 -fmad=false -std=c++17 -o /tmp/host_check_stereo tools/host_check_stereo.cu
//   /tmp/host_check_stereo < pairs.txt   (first line: n, then 9 numbers of F, then n lines x1 y1 x2 y2)
#include <cstdio>
#include <vector>
#include "../object_keypoints_b200/csrc/okp_stereo.cuh"
int main() {
    int n;
    if (scanf("%d", &n) != 1) return 1;
    OkpMat3 F;
    for (int i = 0; i < 9; ++i) if (scanf("%lf", &F.m[i]) != 1) return 1;
    const OkpEpipoles ep = okp_epipoles(F);
    for (int i = 0; i < n; ++i) {
        double x1, y1, x2, y2, o1[2], o2[2];
        if (scanf("%lf %lf %lf %lf", &x1, &y1, &x2, &y2) != 4) return 1;
        okp_correct_pair(F, ep, x1, y1, x2, y2, o1, o2);
        printf("%.17g %.17g %.17g %.17g\n", o1[0], o1[1], o2[0], o2[1]);
    }
    return 0;
}
