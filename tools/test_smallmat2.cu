#include <cstdio>
#include "../include/ll_smallmat.h"
__global__ void __launch_bounds__(32) k(const double* totg, float* matPg, int* flags, float* Tg, double* trace, int iter) {
  __shared__ double tot[28];
  if (threadIdx.x < 28) { double v = 0.0; for (int b = 0; b < 3; ++b) v += totg[b * 28 + threadIdx.x]; tot[threadIdx.x] = v; }
  __syncwarp();
  if (threadIdx.x != 0) return;
  float* T = Tg;
  const int rows = (int)tot[27];
  for (int i = 0; i < 28; ++i) trace[i] = tot[i];
  if (rows < 50) return;
  float AtA[36], AtB[6], A2[36], X[6];
  int k = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) { AtA[r * 6 + c] = AtA[c * 6 + r] = (float)tot[k]; ++k; }
  for (int r = 0; r < 6; ++r) AtB[r] = (float)tot[21 + r];
  for (int i = 0; i < 36; ++i) A2[i] = AtA[i];
  llm::colpiv_qr_solve<6, 6>(A2, AtB, X);
  for (int i = 0; i < 6; ++i) trace[34 + i] = X[i];
  float* matP = matPg;
  if (iter == 0) flags[0] = llm::degeneracy_projector<6>(AtA, 100.f, matP) ? 1 : 0;
  if (flags[0]) {
    float X2[6];
    for (int i = 0; i < 6; ++i) X2[i] = X[i];
    for (int r = 0; r < 6; ++r) { float v = 0.f; for (int c = 0; c < 6; ++c) v += matP[r * 6 + c] * X2[c]; X[r] = v; }
  }
  for (int i = 0; i < 6; ++i) T[i] += X[i];
  for (int i = 0; i < 6; ++i) trace[28 + i] = (double)X[i];
}
int main() {
  double u[21]={5.02468e+04,1.95079e+01,1.64226e+01,8.36225e+00,8.62205e+01,3.12794e+01,1.78237e+03,-1.77829e+02,9.32215e+01,-2.79402e+01,-1.63288e+02,4.97819e+04,-8.29275e+01,-1.28221e+02,2.10462e+01,2.58653e+01,6.58260e-01,-7.75434e+00,1.49597e+03,-2.77888e+00,4.79601e+01};
  double b[6]={-101.24257,-29.56545,153.43656,-2.79859,-30.71684,4.28407};
  double tot[84]={0}; for(int i=0;i<21;i++)tot[i]=u[i]; for(int i=0;i<6;i++)tot[21+i]=b[i]; tot[27]=1610;
  double *dt,*dtr; float *dP,*dT; int* df;
  cudaMalloc(&dt,84*8); cudaMalloc(&dtr,40*8); cudaMalloc(&dP,144); cudaMalloc(&dT,24); cudaMalloc(&df,16);
  cudaMemcpy(dt,tot,84*8,cudaMemcpyHostToDevice); cudaMemset(dP,0,144); cudaMemset(dT,0,24); cudaMemset(df,0,16); cudaMemset(dtr,0,320);
  k<<<1,32>>>(dt,dP,df,dT,dtr,0);
  double tr[40]; cudaError_t e=cudaMemcpy(tr,dtr,320,cudaMemcpyDeviceToHost); int fl; cudaMemcpy(&fl,df,4,cudaMemcpyDeviceToHost);
  printf("err %s flag %d\n",cudaGetErrorString(e),fl);
  printf("X plain: "); for(int i=0;i<6;i++)printf("%g ",tr[34+i]); printf("\nX final: "); for(int i=0;i<6;i++)printf("%g ",tr[28+i]); printf("\n");
}
