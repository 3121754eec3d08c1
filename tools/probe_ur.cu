// Microbenchmarks, round 1 second session: sphere PAIRS as uniform-register operands of FFMA2 (constant bank, LDCU) with ONE
// ray per lane, and 7-op forms of the filter.  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe_ur probe_ur.cu
// Prints one JSON line per probe ("tflops" = 17 flop per test, the bench's algorithmic figure).  Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)
typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c){ float2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return d; }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b){ float2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b){ float2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 bc(float a){ return make_float2(a,a); }
__device__ __forceinline__ float2 neg2(float2 a){ return make_float2(-a.x,-a.y); }

#define NP 544               // spheres (padded), as in the final scene
#define NPAIR (NP/2)
// pair-interleaved constant layout: c_xy[j] = (cx0,cx1,cy0,cy1), c_zk[j] = (cz0,cz1,K0,K1) of sphere pair j
__constant__ float4 c_xy[NPAIR];
__constant__ float4 c_zk[NPAIR];

struct Ray { float dx,dy,dz,mx,my,mz,nod,oo; };
__device__ __forceinline__ Ray make_ray(){
  Ray r; float t=threadIdx.x*0.01f; float ox=13+t, oy=2, oz=3-t; r.dx=-0.9f+t*1e-3f; r.dy=-0.1f+t*0.01f; r.dz=-0.2f-t*1e-3f;
  r.mx=-2*ox; r.my=-2*oy; r.mz=-2*oz; r.nod=-(ox*r.dx+oy*r.dy+oz*r.dz); r.oo=ox*ox+oy*oy+oz*oz; return r; }

// MODE 0: 8-op, all four sphere values from uniform registers (K through FADD2)
// MODE 1: 7-op, K is the addend of the C chain's head (two uniform operands in one FFMA2 unless the compiler moves one)
// MODE 2: 7-op, X/Y/Z from uniform registers, K pairs from shared memory (LDS.128 per two pairs)
// MODE 3: 7-op + threshold compare: FSET + SHF per sphere (disc' >= oo), all from uniform registers, K into the head via register
template<int MODE, int CTAS>
__global__ void __launch_bounds__(256,CTAS) k_ur(float* out, int iters){
  __shared__ float4 s_k[NPAIR/2];
  if (MODE==2){ for(int i=threadIdx.x;i<NPAIR/2;i+=blockDim.x){ float4 a=c_zk[2*i], b=c_zk[2*i+1]; s_k[i]=make_float4(a.z,a.w,b.z,b.w); } __syncthreads(); }
  Ray r=make_ray(); unsigned acc=0;
  for(int it=0; it<iters; ++it){
    for(int w=0; w<NPAIR; w+=16){
      unsigned m=0;
      #pragma unroll
      for(int q=0;q<16;q++){
        const float4 xy=c_xy[w+q], zk=c_zk[w+q];
        const float2 X=make_float2(xy.x,xy.y), Y=make_float2(xy.z,xy.w), Z=make_float2(zk.x,zk.y);
        float2 K=make_float2(zk.z,zk.w);
        if (MODE==2){ const float4 kk=s_k[(w+q)>>1]; K = (q&1)? make_float2(kk.z,kk.w) : make_float2(kk.x,kk.y); }
        const float2 hb=ffma2(X,bc(r.dx),ffma2(Y,bc(r.dy),ffma2(Z,bc(r.dz),bc(r.nod))));
        float2 C0;
        if (MODE==0) C0=fadd2(K,bc(r.oo)); else C0=K;
        const float2 C=ffma2(X,bc(r.mx),ffma2(Y,bc(r.my),ffma2(Z,bc(r.mz),C0)));
        const float2 disc=ffma2(hb,hb,neg2(C));
        if (MODE==3){
          m=__funnelshift_l(disc.x>=r.oo?0x80000000u:0u, m, 1);
          m=__funnelshift_l(disc.y>=r.oo?0x80000000u:0u, m, 1);
        } else {
          m=__funnelshift_l(__float_as_uint(disc.x), m, 1);
          m=__funnelshift_l(__float_as_uint(disc.y), m, 1);
        }
      }
      acc+=__popc(~m); r.oo+=1e-6f; r.nod+=1e-7f;
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

// shared-memory SoA forms (the product's current structure: 4 LDS.128 per 4 spheres), MODE 0: 8-op (baseline), 1: 7-op, 3: 7-op + FSET compare
template<int MODE, int CTAS>
__global__ void __launch_bounds__(256,CTAS) k_smem(const float* __restrict__ g, float* out, int iters){
  __shared__ float4 sm4[NP];
  for(int i=threadIdx.x;i<NP;i+=blockDim.x) sm4[i]=((const float4*)g)[i];
  __syncthreads();
  const int n4=NP/4;
  const float4* CX=sm4; const float4* CY=sm4+n4; const float4* CZ=sm4+2*n4; const float4* KK=sm4+3*n4;
  Ray r=make_ray(); unsigned acc=0;
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n4; w+=8){
      unsigned m=0;
      #pragma unroll
      for(int q=0;q<8;q++){
        const float4 cx=CX[w+q], cy=CY[w+q], cz=CZ[w+q], kk=KK[w+q];
        #pragma unroll
        for(int h=0;h<2;h++){
          const float2 X= h? make_float2(cx.z,cx.w):make_float2(cx.x,cx.y);
          const float2 Y= h? make_float2(cy.z,cy.w):make_float2(cy.x,cy.y);
          const float2 Z= h? make_float2(cz.z,cz.w):make_float2(cz.x,cz.y);
          const float2 K= h? make_float2(kk.z,kk.w):make_float2(kk.x,kk.y);
          const float2 hb=ffma2(X,bc(r.dx),ffma2(Y,bc(r.dy),ffma2(Z,bc(r.dz),bc(r.nod))));
          const float2 C0 = MODE==0 ? fadd2(K,bc(r.oo)) : K;
          const float2 C=ffma2(X,bc(r.mx),ffma2(Y,bc(r.my),ffma2(Z,bc(r.mz),C0)));
          const float2 disc=ffma2(hb,hb,neg2(C));
          if (MODE==3){
            m=__funnelshift_l(disc.x>=r.oo?0x80000000u:0u, m, 1);
            m=__funnelshift_l(disc.y>=r.oo?0x80000000u:0u, m, 1);
          } else {
            m=__funnelshift_l(__float_as_uint(disc.x), m, 1);
            m=__funnelshift_l(__float_as_uint(disc.y), m, 1);
          }
        }
      }
      acc+=__popc(~m); r.oo+=1e-6f; r.nod+=1e-7f;
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b){ float ms; CK(cudaEventElapsedTime(&ms,a,b)); return ms; }

int main(){
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  const int sms=p.multiProcessorCount;
  printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d}\n",p.name,sms);
  std::vector<float> h(4*NP);
  for(int i=0;i<NP;i++){ h[i]=(i%23)-11+0.3f; h[NP+i]=0.2f; h[2*NP+i]=(i/23)-11+0.4f; h[3*NP+i]=0.04f; }
  float* g; CK(cudaMalloc(&g,sizeof(float)*4*NP)); CK(cudaMemcpy(g,h.data(),sizeof(float)*4*NP,cudaMemcpyHostToDevice));
  { std::vector<float4> xy(NPAIR), zk(NPAIR);
    for(int j=0;j<NPAIR;j++){ int a=2*j,b=2*j+1; xy[j]=make_float4(h[a],h[b],h[NP+a],h[NP+b]); zk[j]=make_float4(h[2*NP+a],h[2*NP+b],h[3*NP+a],h[3*NP+b]); }
    CK(cudaMemcpyToSymbol(c_xy,xy.data(),sizeof(float4)*NPAIR)); CK(cudaMemcpyToSymbol(c_zk,zk.data(),sizeof(float4)*NPAIR)); }
  const int threads=256;
  float* out; CK(cudaMalloc(&out,sizeof(float)*threads*sms*8));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int sit=256;
#define RUN(name, kern, ctas_per_sm) do{ const int ctas=sms*(ctas_per_sm); \
    for(int rep=0;rep<2;rep++){ kern<<<ctas,threads>>>(out,sit); } CK(cudaDeviceSynchronize()); \
    float best=1e30f; for(int rep=0;rep<5;rep++){ CK(cudaEventRecord(e0)); kern<<<ctas,threads>>>(out,sit); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms=time_ms(e0,e1); if(ms<best)best=ms; } \
    CK(cudaGetLastError()); double fl=17.0*NP*sit*(double)threads*ctas; \
    printf("{\"probe\":\"%s\",\"ctas_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",name,ctas_per_sm,best,fl/best*1e-9); fflush(stdout); }while(0)
#define RUNS(name, kern, ctas_per_sm) do{ const int ctas=sms*(ctas_per_sm); \
    for(int rep=0;rep<2;rep++){ kern<<<ctas,threads>>>(g,out,sit); } CK(cudaDeviceSynchronize()); \
    float best=1e30f; for(int rep=0;rep<5;rep++){ CK(cudaEventRecord(e0)); kern<<<ctas,threads>>>(g,out,sit); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms=time_ms(e0,e1); if(ms<best)best=ms; } \
    CK(cudaGetLastError()); double fl=17.0*NP*sit*(double)threads*ctas; \
    printf("{\"probe\":\"%s\",\"ctas_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",name,ctas_per_sm,best,fl/best*1e-9); fflush(stdout); }while(0)
  RUNS("smem_8op", (k_smem<0,3>), 3);
  RUNS("smem_7op", (k_smem<1,3>), 3);
  RUNS("smem_7op_fset", (k_smem<3,3>), 3);
  RUN("ur_8op", (k_ur<0,3>), 3);
  RUN("ur_7op_k_uniform", (k_ur<1,3>), 3);
  RUN("ur_7op_k_lds", (k_ur<2,3>), 3);
  RUN("ur_7op_fset", (k_ur<3,3>), 3);
  RUNS("smem_8op", (k_smem<0,4>), 4);
  RUNS("smem_7op", (k_smem<1,4>), 4);
  RUN("ur_8op", (k_ur<0,4>), 4);
  RUN("ur_7op_k_uniform", (k_ur<1,4>), 4);
  RUN("ur_7op_k_lds", (k_ur<2,4>), 4);
  RUN("ur_8op", (k_ur<0,2>), 2);
  RUN("ur_7op_k_lds", (k_ur<2,2>), 2);
  RUN("ur_8op", (k_ur<0,6>), 6);
  RUN("ur_7op_k_lds", (k_ur<2,6>), 6);
  return 0;
}
