#include <cstdio>
__global__ void k(const int* a, const int* b, unsigned* o, int* lo, int* hi) {
  int i = threadIdx.x;
  unsigned z;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(z) : "r"(a[i]), "r"(b[i]));
  o[i] = z; lo[i] = (int)__byte_perm(z, 0, 0x9910); hi[i] = (int)z >> 16;
}
int main() {
  int ha[8] = {1, -1, 40000, -40000, 123, -32768, 32767, 0}, hb[8] = {2, -2, -50000, 50000, -456, 7, -7, 65536};
  int *a, *b, *lo, *hi; unsigned* o;
  cudaMalloc(&a, 32); cudaMalloc(&b, 32); cudaMalloc(&o, 32); cudaMalloc(&lo, 32); cudaMalloc(&hi, 32);
  cudaMemcpy(a, ha, 32, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, 32, cudaMemcpyHostToDevice);
  k<<<1, 8>>>(a, b, o, lo, hi);
  unsigned ho[8]; int hlo[8], hhi[8];
  cudaMemcpy(ho, o, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hlo, lo, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hhi, hi, 32, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 8; i++) printf("a=%d b=%d -> %08x lo=%d hi=%d\n", ha[i], hb[i], ho[i], hlo[i], hhi[i]);
}
