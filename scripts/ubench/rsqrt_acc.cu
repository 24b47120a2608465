#include <cstdio>
#include <cmath>
__device__ double f2(double a){ double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); const double h=0.5*a; y=fma(y,fma(-h*y,y,0.5),y); y=fma(y,fma(-h*y,y,0.5),y); return y;}
__device__ double f3(double a){ double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); const double t=a*y; const double e=fma(-t,y,1.0); const double ye=y*e; const double p=fma(0.375,e,0.5); return fma(ye,p,y);}
__device__ double f0(double a){ double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); return y;}
__global__ void k(double* out){ double m2=0,m3=0,m0=0; for(int i=threadIdx.x;i<4000000;i+=blockDim.x){ double a=exp2((i%200)-100.)*(1.0+ (i*0.61803398875-floor(i*0.61803398875))); double ex=1.0/sqrt(a); m2=fmax(m2,fabs(f2(a)-ex)/ex); m3=fmax(m3,fabs(f3(a)-ex)/ex); m0=fmax(m0,fabs(f0(a)-ex)/ex);} 
 __shared__ double s[3][256]; s[0][threadIdx.x]=m2; s[1][threadIdx.x]=m3; s[2][threadIdx.x]=m0; __syncthreads(); if(threadIdx.x==0){ for(int q=0;q<3;++q){double m=0; for(int i=0;i<256;++i) m=fmax(m,s[q][i]); out[q]=m;} } }
int main(){ double* d; cudaMalloc(&d,24); k<<<1,256>>>(d); double h[3]; cudaMemcpy(h,d,24,cudaMemcpyDeviceToHost); printf("max rel err: two Newton %.3e  one cubic %.3e  raw approx %.3e (%s)\n",h[0],h[1],h[2],cudaGetErrorString(cudaGetLastError())); }
