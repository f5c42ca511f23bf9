// developer tool: run ll_smallmat.h on the device and on the host on the same 6x6 system
#include <cstdio>
#include "../include/ll_smallmat.h"
__host__ __device__ void run(const float* AtA, const float* AtB, float* out) {
  float A2[36], X[6], P[36], ev[6], V[36];
  for (int i = 0; i < 36; ++i) A2[i] = AtA[i];
  llm::colpiv_qr_solve<6, 6>(A2, AtB, X);
  for (int i = 0; i < 6; ++i) out[i] = X[i];
  bool deg = llm::degeneracy_projector<6>(AtA, 100.f, P);
  out[6] = deg ? 1.f : 0.f;
  llm::self_adjoint_eigen<6>(AtA, ev, V);
  for (int i = 0; i < 6; ++i) out[7 + i] = ev[i];
}
__global__ void k(const float* AtA, const float* AtB, float* out) { run(AtA, AtB, out); }
int main() {
  double u[21]={5.02468e+04,1.95079e+01,1.64226e+01,8.36225e+00,8.62205e+01,3.12794e+01,1.78237e+03,-1.77829e+02,9.32215e+01,-2.79402e+01,-1.63288e+02,4.97819e+04,-8.29275e+01,-1.28221e+02,2.10462e+01,2.58653e+01,6.58260e-01,-7.75434e+00,1.49597e+03,-2.77888e+00,4.79601e+01};
  float AtA[36]; int kk=0; for(int r=0;r<6;r++)for(int c=r;c<6;c++){AtA[r*6+c]=AtA[c*6+r]=(float)u[kk++];}
  float AtB[6]={-101.24257f,-29.56545f,153.43656f,-2.79859f,-30.71684f,4.28407f};
  float ho[13]; run(AtA, AtB, ho);
  float *dA,*dB,*dO; cudaMalloc(&dA,144); cudaMalloc(&dB,24); cudaMalloc(&dO,52);
  cudaMemcpy(dA,AtA,144,cudaMemcpyHostToDevice); cudaMemcpy(dB,AtB,24,cudaMemcpyHostToDevice);
  k<<<1,1>>>(dA,dB,dO); float go[13]; cudaError_t e = cudaMemcpy(go,dO,52,cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(e));
  printf("host: "); for(int i=0;i<13;i++)printf("%g ",ho[i]); printf("\n");
  printf("dev : "); for(int i=0;i<13;i++)printf("%g ",go[i]); printf("\n");
  return 0;
}
