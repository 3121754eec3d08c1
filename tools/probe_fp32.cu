// Microbenchmarks that size the sphere-scan inner loop on B200 (sm_100a).
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe_fp32 probe_fp32.cu
// Prints one JSON line per probe. Used to choose the scan variant; not part of the product path.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c){ float2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return d; }
__device__ __forceinline__ float2 fsub2(float2 a, float2 b){ float2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b){ float2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 neg2(float2 a){ return make_float2(-a.x,-a.y); }
__device__ __forceinline__ float2 bc(float a){ return make_float2(a,a); }

// ---- A: scalar FFMA peak -------------------------------------------------
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b){
  float x[8];
  #pragma unroll
  for(int i=0;i<8;i++) x[i]=threadIdx.x*0.001f+i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int r=0;r<8;r++){
      #pragma unroll
      for(int i=0;i<8;i++) x[i]=fmaf(x[i],a,b);
    }
  }
  float s=0; for(int i=0;i<8;i++) s+=x[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// ---- B: packed FFMA2 peak -------------------------------------------------
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b){
  float2 x[8];
  #pragma unroll
  for(int i=0;i<8;i++) x[i]=make_float2(threadIdx.x*0.001f+i, i*0.5f);
  float2 A=bc(a), B=bc(b);
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int r=0;r<8;r++){
      #pragma unroll
      for(int i=0;i<8;i++) x[i]=ffma2(x[i],A,B);
    }
  }
  float s=0; for(int i=0;i<8;i++) s+=x[i].x+x[i].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// ---- C: FFMA2 with one SHF per FFMA2 interleaved (ALU co-issue) ---------------
__global__ void __launch_bounds__(256) k_ffma2_shf(float* out, int iters, float a, float b){
  float2 x[8]; unsigned m=threadIdx.x;
  #pragma unroll
  for(int i=0;i<8;i++) x[i]=make_float2(threadIdx.x*0.001f+i, i*0.5f);
  float2 A=bc(a), B=bc(b);
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int r=0;r<8;r++){
      #pragma unroll
      for(int i=0;i<8;i++){ x[i]=ffma2(x[i],A,B); }
      #pragma unroll
      for(int i=0;i<8;i+=2){ m=__funnelshift_l(__float_as_uint(x[i].x), m, 1); }
    }
  }
  float s=m; for(int i=0;i<8;i++) s+=x[i].x+x[i].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// ---- D: LDS.128 broadcast throughput -----------------------------------------
__global__ void __launch_bounds__(256) k_lds128(float* out, int iters){
  __shared__ float4 sm[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) sm[i]=make_float4(i,i+1,i+2,i+3);
  __syncthreads();
  float4 acc=make_float4(0,0,0,0);
  for(int it=0; it<iters; ++it){
    #pragma unroll 16
    for(int i=0;i<1024;i++){ float4 v=sm[i]; acc.x+=v.x; acc.y+=v.y; acc.z+=v.z; acc.w+=v.w; }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc.x+acc.y+acc.z+acc.w;
}
// ---- E1: scan from shared SoA, R rays per thread, packed over sphere pairs ------
// filter: disc = hb^2 - |oc|^2 + r2 ; sign bit of each result funnel-shifted into a word
template<int R>
__global__ void __launch_bounds__(256) k_scan_smem(const float* __restrict__ g, float* out, int n, int iters){
  extern __shared__ float4 sm4[];   // [4][n/4] : cx | cy | cz | r2
  const int n4=n/4;
  for(int i=threadIdx.x;i<n;i+=blockDim.x) ((float*)sm4)[i]=g[i];
  __syncthreads();
  const float4* CX=sm4; const float4* CY=sm4+n4; const float4* CZ=sm4+2*n4; const float4* RR=sm4+3*n4;
  float ox[R],oy[R],oz[R],dx[R],dy[R],dz[R]; unsigned acc[R];
  #pragma unroll
  for(int r=0;r<R;r++){ float t=(threadIdx.x*R+r)*0.01f; ox[r]=13+t; oy[r]=2; oz[r]=3-t; dx[r]=-0.9f+t*1e-3f; dy[r]=-0.1f+t*0.01f; dz[r]=-0.2f-t*1e-3f; acc[r]=0; }
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n4; w+=8){           // 32 spheres per word
      unsigned m[R];
      #pragma unroll
      for(int r=0;r<R;r++) m[r]=0;
      #pragma unroll
      for(int q=0;q<8;q++){
        float4 cx=CX[w+q], cy=CY[w+q], cz=CZ[w+q], r2=RR[w+q];
        #pragma unroll
        for(int h=0;h<2;h++){
          float2 X= h? make_float2(cx.z,cx.w):make_float2(cx.x,cx.y);
          float2 Y= h? make_float2(cy.z,cy.w):make_float2(cy.x,cy.y);
          float2 Z= h? make_float2(cz.z,cz.w):make_float2(cz.x,cz.y);
          float2 Q= h? make_float2(r2.z,r2.w):make_float2(r2.x,r2.y);
          #pragma unroll
          for(int r=0;r<R;r++){
            float2 ocx=fsub2(X,bc(ox[r])), ocy=fsub2(Y,bc(oy[r])), ocz=fsub2(Z,bc(oz[r]));
            float2 hb=ffma2(ocz,bc(dz[r]),ffma2(ocy,bc(dy[r]),fmul2(ocx,bc(dx[r]))));
            float2 nc=ffma2(neg2(ocx),ocx,ffma2(neg2(ocy),ocy,ffma2(neg2(ocz),ocz,Q)));
            float2 disc=ffma2(hb,hb,nc);
            m[r]=__funnelshift_l(__float_as_uint(disc.x), m[r], 1);
            m[r]=__funnelshift_l(__float_as_uint(disc.y), m[r], 1);
          }
        }
      }
      #pragma unroll
      for(int r=0;r<R;r++){ acc[r]+=__popc(~m[r]); ox[r]+=1e-6f; }
    }
  }
  unsigned s=0;
  #pragma unroll
  for(int r=0;r<R;r++) s+=acc[r];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// ---- E0: same scan, scalar FFMA (no packing), R rays per thread -------------------
template<int R>
__global__ void __launch_bounds__(256) k_scan_smem_scalar(const float* __restrict__ g, float* out, int n, int iters){
  extern __shared__ float4 sm4[];   // AoS: (cx,cy,cz,r2) per sphere
  for(int i=threadIdx.x;i<n;i+=blockDim.x) sm4[i]=make_float4(g[i],g[n+i],g[2*n+i],g[3*n+i]);
  __syncthreads();
  float ox[R],oy[R],oz[R],dx[R],dy[R],dz[R]; unsigned acc[R];
  #pragma unroll
  for(int r=0;r<R;r++){ float t=(threadIdx.x*R+r)*0.01f; ox[r]=13+t; oy[r]=2; oz[r]=3-t; dx[r]=-0.9f+t*1e-3f; dy[r]=-0.1f+t*0.01f; dz[r]=-0.2f-t*1e-3f; acc[r]=0; }
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n; w+=32){
      unsigned m[R];
      #pragma unroll
      for(int r=0;r<R;r++) m[r]=0;
      #pragma unroll
      for(int q=0;q<32;q++){
        float4 s=sm4[w+q];
        #pragma unroll
        for(int r=0;r<R;r++){
          float ocx=s.x-ox[r], ocy=s.y-oy[r], ocz=s.z-oz[r];
          float hb=fmaf(ocz,dz[r],fmaf(ocy,dy[r],ocx*dx[r]));
          float nc=fmaf(-ocx,ocx,fmaf(-ocy,ocy,fmaf(-ocz,ocz,s.w)));
          float disc=fmaf(hb,hb,nc);
          m[r]=__funnelshift_l(__float_as_uint(disc), m[r], 1);
        }
      }
      #pragma unroll
      for(int r=0;r<R;r++){ acc[r]+=__popc(~m[r]); ox[r]+=1e-6f; }
    }
  }
  unsigned s=0;
  #pragma unroll
  for(int r=0;r<R;r++) s+=acc[r];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// ---- E2: scan from the constant bank, 2 rays per lane packed in the halves --------
#define CN 4096
__constant__ float4 c_sph[CN];   // (cx,cy,cz,r2)
template<int RP>   // RP ray PAIRS per thread
__global__ void __launch_bounds__(256) k_scan_const(float* out, int n, int iters){
  float2 ox[RP],oy[RP],oz[RP],dx[RP],dy[RP],dz[RP]; unsigned acc=0;
  #pragma unroll
  for(int r=0;r<RP;r++){ float t=(threadIdx.x*RP+r)*0.01f; ox[r]=make_float2(13+t,13-t); oy[r]=bc(2); oz[r]=make_float2(3-t,3+t); dx[r]=make_float2(-0.9f+t*1e-3f,-0.9f-t*1e-3f); dy[r]=make_float2(-0.1f+t*0.01f,-0.1f); dz[r]=make_float2(-0.2f-t*1e-3f,-0.2f+t*1e-3f); }
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n; w+=32){
      unsigned m0[RP], m1[RP];
      #pragma unroll
      for(int r=0;r<RP;r++){ m0[r]=0; m1[r]=0; }
      #pragma unroll
      for(int q=0;q<32;q++){
        float4 s=c_sph[w+q];
        #pragma unroll
        for(int r=0;r<RP;r++){
          float2 ocx=fsub2(bc(s.x),ox[r]), ocy=fsub2(bc(s.y),oy[r]), ocz=fsub2(bc(s.z),oz[r]);
          float2 hb=ffma2(ocz,dz[r],ffma2(ocy,dy[r],fmul2(ocx,dx[r])));
          float2 nc=ffma2(neg2(ocx),ocx,ffma2(neg2(ocy),ocy,ffma2(neg2(ocz),ocz,bc(s.w))));
          float2 disc=ffma2(hb,hb,nc);
          m0[r]=__funnelshift_l(__float_as_uint(disc.x), m0[r], 1);
          m1[r]=__funnelshift_l(__float_as_uint(disc.y), m1[r], 1);
        }
      }
      #pragma unroll
      for(int r=0;r<RP;r++){ acc+=__popc(~m0[r])+__popc(~m1[r]); ox[r].x+=1e-6f; }
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}


// ---- E3: 8-op expanded filter: hb = c.d - o.d ; C = K + |o|^2 - 2 c.o ; disc = hb^2 - C ---------------
__device__ __forceinline__ float2 fadd2(float2 a, float2 b){ float2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
template<int R>
__global__ void __launch_bounds__(256) k_scan_smem8(const float* __restrict__ g, float* out, int n, int iters){
  extern __shared__ float4 sm4[];
  const int n4=n/4;
  for(int i=threadIdx.x;i<n;i+=blockDim.x) ((float*)sm4)[i]=g[i];
  __syncthreads();
  const float4* CX=sm4; const float4* CY=sm4+n4; const float4* CZ=sm4+2*n4; const float4* KK=sm4+3*n4;
  float m2ox[R],m2oy[R],m2oz[R],dx[R],dy[R],dz[R],nod[R],oo[R]; unsigned acc[R];
  #pragma unroll
  for(int r=0;r<R;r++){ float t=(threadIdx.x*R+r)*0.01f; float ox=13+t, oy=2, oz=3-t; dx[r]=-0.9f+t*1e-3f; dy[r]=-0.1f+t*0.01f; dz[r]=-0.2f-t*1e-3f;
     m2ox[r]=-2*ox; m2oy[r]=-2*oy; m2oz[r]=-2*oz; nod[r]=-(ox*dx[r]+oy*dy[r]+oz*dz[r]); oo[r]=ox*ox+oy*oy+oz*oz; acc[r]=0; }
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n4; w+=8){
      unsigned m[R];
      #pragma unroll
      for(int r=0;r<R;r++) m[r]=0;
      #pragma unroll
      for(int q=0;q<8;q++){
        float4 cx=CX[w+q], cy=CY[w+q], cz=CZ[w+q], kk=KK[w+q];
        #pragma unroll
        for(int h=0;h<2;h++){
          float2 X= h? make_float2(cx.z,cx.w):make_float2(cx.x,cx.y);
          float2 Y= h? make_float2(cy.z,cy.w):make_float2(cy.x,cy.y);
          float2 Z= h? make_float2(cz.z,cz.w):make_float2(cz.x,cz.y);
          float2 K= h? make_float2(kk.z,kk.w):make_float2(kk.x,kk.y);
          #pragma unroll
          for(int r=0;r<R;r++){
            float2 hb=ffma2(X,bc(dx[r]),ffma2(Y,bc(dy[r]),ffma2(Z,bc(dz[r]),bc(nod[r]))));
            float2 C=ffma2(X,bc(m2ox[r]),ffma2(Y,bc(m2oy[r]),ffma2(Z,bc(m2oz[r]),fadd2(K,bc(oo[r])))));
            float2 disc=ffma2(hb,hb,neg2(C));
            m[r]=__funnelshift_l(__float_as_uint(disc.x), m[r], 1);
            m[r]=__funnelshift_l(__float_as_uint(disc.y), m[r], 1);
          }
        }
      }
      #pragma unroll
      for(int r=0;r<R;r++){ acc[r]+=__popc(~m[r]); oo[r]+=1e-6f; }
    }
  }
  unsigned s=0;
  #pragma unroll
  for(int r=0;r<R;r++) s+=acc[r];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// ---- E4: 8-op filter with explicit software prefetch of the next sphere quads (register double-buffering) ----
template<int PF>
__global__ void __launch_bounds__(256,3) k_scan_smem8_pf(const float* __restrict__ g, float* out, int n, int iters){
  extern __shared__ float4 sm4[];
  const int n4=n/4;
  for(int i=threadIdx.x;i<n;i+=blockDim.x) ((float*)sm4)[i]=g[i];
  __syncthreads();
  const float4* CX=sm4; const float4* CY=sm4+n4; const float4* CZ=sm4+2*n4; const float4* KK=sm4+3*n4;
  float t=(threadIdx.x)*0.01f; float ox=13+t, oy=2, oz=3-t; float dx=-0.9f+t*1e-3f, dy=-0.1f+t*0.01f, dz=-0.2f-t*1e-3f;
  float m2ox=-2*ox, m2oy=-2*oy, m2oz=-2*oz, nod=-(ox*dx+oy*dy+oz*dz), oo=ox*ox+oy*oy+oz*oz; unsigned acc=0;
  for(int it=0; it<iters; ++it){
    float4 bx[PF+1],by[PF+1],bz[PF+1],bk[PF+1];
    #pragma unroll
    for(int p=0;p<PF;p++){ bx[p]=CX[p]; by[p]=CY[p]; bz[p]=CZ[p]; bk[p]=KK[p]; }
    for(int w=0; w<n4; w+=8){
      unsigned m=0;
      #pragma unroll
      for(int q=0;q<8;q++){
        // prefetch quad (w+q+PF) into slot (q+PF)%(PF+1); wraps harmlessly at the end (index clamped)
        int nx = w+q+PF; nx = nx < n4 ? nx : n4-1;
        bx[(q+PF)%(PF+1)]=CX[nx]; by[(q+PF)%(PF+1)]=CY[nx]; bz[(q+PF)%(PF+1)]=CZ[nx]; bk[(q+PF)%(PF+1)]=KK[nx];
        float4 cx=bx[q%(PF+1)], cy=by[q%(PF+1)], cz=bz[q%(PF+1)], kk=bk[q%(PF+1)];
        #pragma unroll
        for(int h=0;h<2;h++){
          float2 X= h? make_float2(cx.z,cx.w):make_float2(cx.x,cx.y);
          float2 Y= h? make_float2(cy.z,cy.w):make_float2(cy.x,cy.y);
          float2 Z= h? make_float2(cz.z,cz.w):make_float2(cz.x,cz.y);
          float2 K= h? make_float2(kk.z,kk.w):make_float2(kk.x,kk.y);
          float2 hb=ffma2(X,bc(dx),ffma2(Y,bc(dy),ffma2(Z,bc(dz),bc(nod))));
          float2 C=ffma2(X,bc(m2ox),ffma2(Y,bc(m2oy),ffma2(Z,bc(m2oz),fadd2(K,bc(oo)))));
          float2 disc=ffma2(hb,hb,neg2(C));
          m=__funnelshift_l(__float_as_uint(disc.x), m, 1);
          m=__funnelshift_l(__float_as_uint(disc.y), m, 1);
        }
      }
      acc+=__popc(~m); oo+=1e-6f;
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

// ---- E5: ablations of the 8-op scan: MODE 0 = full, 1 = no SHF (xor-accumulate one in 8 results), 2 = no LDS (sphere quads held in registers), 3 = neither
template<int MODE>
__global__ void __launch_bounds__(256) k_scan8_abl(const float* __restrict__ g, float* out, int n, int iters){
  extern __shared__ float4 sm4[];
  const int n4=n/4;
  for(int i=threadIdx.x;i<n;i+=blockDim.x) ((float*)sm4)[i]=g[i];
  __syncthreads();
  const float4* CX=sm4; const float4* CY=sm4+n4; const float4* CZ=sm4+2*n4; const float4* KK=sm4+3*n4;
  float t=(threadIdx.x)*0.01f; float ox=13+t, oy=2, oz=3-t; float dx=-0.9f+t*1e-3f, dy=-0.1f+t*0.01f, dz=-0.2f-t*1e-3f;
  float m2ox=-2*ox, m2oy=-2*oy, m2oz=-2*oz, nod=-(ox*dx+oy*dy+oz*dz), oo=ox*ox+oy*oy+oz*oz; unsigned acc=0;
  float4 rx=CX[threadIdx.x&7], ry=CY[threadIdx.x&7], rz=CZ[threadIdx.x&7], rk=KK[threadIdx.x&7];
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n4; w+=8){
      unsigned m=0;
      #pragma unroll
      for(int q=0;q<8;q++){
        float4 cx,cy,cz,kk;
        if (MODE&2) { cx=rx; cy=ry; cz=rz; kk=rk; rx.x+=1e-7f; } else { cx=CX[w+q]; cy=CY[w+q]; cz=CZ[w+q]; kk=KK[w+q]; }
        #pragma unroll
        for(int h=0;h<2;h++){
          float2 X= h? make_float2(cx.z,cx.w):make_float2(cx.x,cx.y);
          float2 Y= h? make_float2(cy.z,cy.w):make_float2(cy.x,cy.y);
          float2 Z= h? make_float2(cz.z,cz.w):make_float2(cz.x,cz.y);
          float2 K= h? make_float2(kk.z,kk.w):make_float2(kk.x,kk.y);
          float2 hb=ffma2(X,bc(dx),ffma2(Y,bc(dy),ffma2(Z,bc(dz),bc(nod))));
          float2 C=ffma2(X,bc(m2ox),ffma2(Y,bc(m2oy),ffma2(Z,bc(m2oz),fadd2(K,bc(oo)))));
          float2 disc=ffma2(hb,hb,neg2(C));
          if (MODE&1) { if (h==1 && (q&3)==3) m^=__float_as_uint(disc.x)^__float_as_uint(disc.y); else { oo+=disc.x*0.f; } }
          else { m=__funnelshift_l(__float_as_uint(disc.x), m, 1); m=__funnelshift_l(__float_as_uint(disc.y), m, 1); }
        }
      }
      acc+=__popc(~m); oo+=1e-6f;
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

// ---- E6: the 8-op scan with NO shared-memory loads: sphere quads live in registers and are perturbed with one
//      LOP3 per register per use so that nothing is loop-invariant.  Isolates the LDS cost.
__global__ void __launch_bounds__(256) k_scan8_regs(const float* __restrict__ g, float* out, int n, int iters){
  const int n4=n/4;
  float t=(threadIdx.x)*0.01f; float ox=13+t, oy=2, oz=3-t; float dx=-0.9f+t*1e-3f, dy=-0.1f+t*0.01f, dz=-0.2f-t*1e-3f;
  float m2ox=-2*ox, m2oy=-2*oy, m2oz=-2*oz, nod=-(ox*dx+oy*dy+oz*dz), oo=ox*ox+oy*oy+oz*oz; unsigned acc=0;
  const float4* G=(const float4*)g;
  float4 rx=G[threadIdx.x&7], ry=G[n4+(threadIdx.x&7)], rz=G[2*n4+(threadIdx.x&7)], rk=G[3*n4+(threadIdx.x&7)];
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n4; w+=8){
      unsigned m=0;
      #pragma unroll
      for(int q=0;q<8;q++){
        unsigned j=(unsigned)(w+q)&3u;     // flips low mantissa bits only
        float4 cx=make_float4(__uint_as_float(__float_as_uint(rx.x)^j),__uint_as_float(__float_as_uint(rx.y)^j),__uint_as_float(__float_as_uint(rx.z)^j),__uint_as_float(__float_as_uint(rx.w)^j));
        float4 cy=make_float4(__uint_as_float(__float_as_uint(ry.x)^j),__uint_as_float(__float_as_uint(ry.y)^j),__uint_as_float(__float_as_uint(ry.z)^j),__uint_as_float(__float_as_uint(ry.w)^j));
        float4 cz=make_float4(__uint_as_float(__float_as_uint(rz.x)^j),__uint_as_float(__float_as_uint(rz.y)^j),__uint_as_float(__float_as_uint(rz.z)^j),__uint_as_float(__float_as_uint(rz.w)^j));
        float4 kk=make_float4(__uint_as_float(__float_as_uint(rk.x)^j),__uint_as_float(__float_as_uint(rk.y)^j),__uint_as_float(__float_as_uint(rk.z)^j),__uint_as_float(__float_as_uint(rk.w)^j));
        #pragma unroll
        for(int h=0;h<2;h++){
          float2 X= h? make_float2(cx.z,cx.w):make_float2(cx.x,cx.y);
          float2 Y= h? make_float2(cy.z,cy.w):make_float2(cy.x,cy.y);
          float2 Z= h? make_float2(cz.z,cz.w):make_float2(cz.x,cz.y);
          float2 K= h? make_float2(kk.z,kk.w):make_float2(kk.x,kk.y);
          float2 hb=ffma2(X,bc(dx),ffma2(Y,bc(dy),ffma2(Z,bc(dz),bc(nod))));
          float2 C=ffma2(X,bc(m2ox),ffma2(Y,bc(m2oy),ffma2(Z,bc(m2oz),fadd2(K,bc(oo)))));
          float2 disc=ffma2(hb,hb,neg2(C));
          m=__funnelshift_l(__float_as_uint(disc.x), m, 1); m=__funnelshift_l(__float_as_uint(disc.y), m, 1);
        }
      }
      acc+=__popc(~m); oo+=1e-6f;
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

// ---- E7: 8-op scan, op-major order over NP sphere pairs (operand-reuse friendly: consecutive FFMA2 share the scalar operand)
template<int NP>   // pairs per step: 2 (one quad), 4 (two quads)
__global__ void __launch_bounds__(256,3) k_scan8_opmajor(const float* __restrict__ g, float* out, int n, int iters){
  extern __shared__ float4 sm4[];
  const int n4=n/4;
  for(int i=threadIdx.x;i<n;i+=blockDim.x) ((float*)sm4)[i]=g[i];
  __syncthreads();
  const float4* CX=sm4; const float4* CY=sm4+n4; const float4* CZ=sm4+2*n4; const float4* KK=sm4+3*n4;
  float t=(threadIdx.x)*0.01f; float ox=13+t, oy=2, oz=3-t; float dx=-0.9f+t*1e-3f, dy=-0.1f+t*0.01f, dz=-0.2f-t*1e-3f;
  float m2ox=-2*ox, m2oy=-2*oy, m2oz=-2*oz, nod=-(ox*dx+oy*dy+oz*dz), oo=ox*ox+oy*oy+oz*oz; unsigned acc=0;
  constexpr int NQ=NP/2;
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n4; w+=8){
      unsigned m=0;
      #pragma unroll
      for(int q=0;q<8;q+=NQ){
        float2 X[NP],Y[NP],Z[NP],K[NP],hb[NP],C[NP];
        #pragma unroll
        for(int j=0;j<NQ;j++){ float4 cx=CX[w+q+j], cy=CY[w+q+j], cz=CZ[w+q+j], kk=KK[w+q+j];
          X[2*j]=make_float2(cx.x,cx.y); X[2*j+1]=make_float2(cx.z,cx.w); Y[2*j]=make_float2(cy.x,cy.y); Y[2*j+1]=make_float2(cy.z,cy.w);
          Z[2*j]=make_float2(cz.x,cz.y); Z[2*j+1]=make_float2(cz.z,cz.w); K[2*j]=make_float2(kk.x,kk.y); K[2*j+1]=make_float2(kk.z,kk.w); }
        #pragma unroll
        for(int p=0;p<NP;p++) C[p]=fadd2(K[p],bc(oo));
        #pragma unroll
        for(int p=0;p<NP;p++) hb[p]=ffma2(Z[p],bc(dz),bc(nod));
        #pragma unroll
        for(int p=0;p<NP;p++) C[p]=ffma2(Z[p],bc(m2oz),C[p]);
        #pragma unroll
        for(int p=0;p<NP;p++) hb[p]=ffma2(Y[p],bc(dy),hb[p]);
        #pragma unroll
        for(int p=0;p<NP;p++) C[p]=ffma2(Y[p],bc(m2oy),C[p]);
        #pragma unroll
        for(int p=0;p<NP;p++) hb[p]=ffma2(X[p],bc(dx),hb[p]);
        #pragma unroll
        for(int p=0;p<NP;p++) C[p]=ffma2(X[p],bc(m2ox),C[p]);
        #pragma unroll
        for(int p=0;p<NP;p++){ float2 disc=ffma2(hb[p],hb[p],neg2(C[p]));
          m=__funnelshift_l(__float_as_uint(disc.x), m, 1); m=__funnelshift_l(__float_as_uint(disc.y), m, 1); }
      }
      acc+=__popc(~m); oo+=1e-6f;
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

// ---- F: FFMA2 operand-form probes (register-file bandwidth): acc[i] = fma(X[i], s, acc[i]) with
//      FORM 0: X pair, s pair (same for all), acc pair      FORM 1: X pair, s scalar-broadcast (4 different), acc pair
//      FORM 2: like 1 but the addend is a scalar-broadcast constant for half of the ops (as in the scan's chain heads)
template<int FORM>
__global__ void __launch_bounds__(256) k_ffma2_forms(float* out, int iters, float a, float b){
  float2 X[8], acc[8]; float sc[4];
  #pragma unroll
  for(int i=0;i<8;i++){ X[i]=make_float2(threadIdx.x*0.001f+i, i*0.5f+a); acc[i]=make_float2(i*b, i+a); }
  #pragma unroll
  for(int i=0;i<4;i++) sc[i]=a*(i+1)+threadIdx.x*1e-6f;
  float2 S=make_float2(a,a*1.5f);
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int r=0;r<8;r++){
      #pragma unroll
      for(int i=0;i<8;i++){
        if (FORM==0) acc[i]=ffma2(X[i],S,acc[i]);
        else if (FORM==1) acc[i]=ffma2(X[i],bc(sc[(i+r)&3]),acc[i]);
        else { if ((i+r)&1) acc[i]=ffma2(X[i],bc(sc[(i+r)&3]),acc[i]); else acc[i]=ffma2(X[(i+1)&7],bc(sc[(i+r)&3]),ffma2(acc[i],bc(0.f),bc(sc[r&3]))); }
      }
    }
    X[it&7].x+=1e-7f;
  }
  float s=0; for(int i=0;i<8;i++) s+=acc[i].x+acc[i].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// ---- E8: 8-op scan, sphere scalars from the CONSTANT BANK (uniform registers), two rays per lane in the f32x2 halves.
//      Every FFMA2 then has at most two register-file pair operands (the third is a uniform-register broadcast).
template<int RP>
__global__ void __launch_bounds__(256) k_scan8_const(float* out, int n, int iters){
  float2 dx[RP],dy[RP],dz[RP],m2ox[RP],m2oy[RP],m2oz[RP],nod[RP],oo[RP]; unsigned acc=0;
  #pragma unroll
  for(int r=0;r<RP;r++){ float t=(threadIdx.x*RP+r)*0.01f; float2 ox=make_float2(13+t,13-t), oy=make_float2(2,2.1f), oz=make_float2(3-t,3+t);
    dx[r]=make_float2(-0.9f+t*1e-3f,-0.9f-t*1e-3f); dy[r]=make_float2(-0.1f+t*0.01f,-0.1f); dz[r]=make_float2(-0.2f-t*1e-3f,-0.2f+t*1e-3f);
    m2ox[r]=make_float2(-2*ox.x,-2*ox.y); m2oy[r]=make_float2(-2*oy.x,-2*oy.y); m2oz[r]=make_float2(-2*oz.x,-2*oz.y);
    nod[r]=make_float2(-(ox.x*dx[r].x+oy.x*dy[r].x+oz.x*dz[r].x),-(ox.y*dx[r].y+oy.y*dy[r].y+oz.y*dz[r].y));
    oo[r]=make_float2(ox.x*ox.x+oy.x*oy.x+oz.x*oz.x, ox.y*ox.y+oy.y*oy.y+oz.y*oz.y); }
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n; w+=32){
      unsigned m0[RP], m1[RP];
      #pragma unroll
      for(int r=0;r<RP;r++){ m0[r]=0; m1[r]=0; }
      #pragma unroll
      for(int q=0;q<32;q++){
        float4 s=c_sph[w+q];     // (cx,cy,cz,K) warp-uniform
        #pragma unroll
        for(int r=0;r<RP;r++){
          float2 hb=ffma2(bc(s.x),dx[r],ffma2(bc(s.y),dy[r],ffma2(bc(s.z),dz[r],nod[r])));
          float2 C=ffma2(bc(s.x),m2ox[r],ffma2(bc(s.y),m2oy[r],ffma2(bc(s.z),m2oz[r],fadd2(bc(s.w),oo[r]))));
          float2 disc=ffma2(hb,hb,neg2(C));
          m0[r]=__funnelshift_l(__float_as_uint(disc.x), m0[r], 1);
          m1[r]=__funnelshift_l(__float_as_uint(disc.y), m1[r], 1);
        }
      }
      #pragma unroll
      for(int r=0;r<RP;r++){ acc+=__popc(~m0[r])+__popc(~m1[r]); oo[r].x+=1e-6f; }
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

// ---- E9: is the constant-bank variant limited by the uniform loads?  Each loaded sphere is used for REUSE ray pairs
//      (RP = REUSE here, but only one LDCU per sphere): if tests/s per FFMA2 rises with REUSE, LDCU was the limiter.
template<int RP>
__global__ void __launch_bounds__(256) k_scan8_const_ur(float* out, int n, int iters){
  float2 dx[RP],dy[RP],dz[RP],m2ox[RP],m2oy[RP],m2oz[RP],nod[RP],oo[RP]; unsigned acc=0;
  #pragma unroll
  for(int r=0;r<RP;r++){ float t=(threadIdx.x*RP+r)*0.01f; float2 ox=make_float2(13+t,13-t), oy=make_float2(2,2.1f), oz=make_float2(3-t,3+t);
    dx[r]=make_float2(-0.9f+t*1e-3f,-0.9f-t*1e-3f); dy[r]=make_float2(-0.1f+t*0.01f,-0.1f); dz[r]=make_float2(-0.2f-t*1e-3f,-0.2f+t*1e-3f);
    m2ox[r]=make_float2(-2*ox.x,-2*ox.y); m2oy[r]=make_float2(-2*oy.x,-2*oy.y); m2oz[r]=make_float2(-2*oz.x,-2*oz.y);
    nod[r]=make_float2(-(ox.x*dx[r].x+oy.x*dy[r].x+oz.x*dz[r].x),-(ox.y*dx[r].y+oy.y*dy[r].y+oz.y*dz[r].y));
    oo[r]=make_float2(ox.x*ox.x+oy.x*oy.x+oz.x*oz.x, ox.y*ox.y+oy.y*oy.y+oz.y*oz.y); }
  for(int it=0; it<iters; ++it){
    for(int w=0; w<n; w+=16){
      unsigned m0[RP], m1[RP];
      #pragma unroll
      for(int r=0;r<RP;r++){ m0[r]=0; m1[r]=0; }
      #pragma unroll
      for(int q=0;q<16;q++){
        float4 s=c_sph[w+q];
        #pragma unroll
        for(int r=0;r<RP;r++){
          float2 hb=ffma2(bc(s.x),dx[r],ffma2(bc(s.y),dy[r],ffma2(bc(s.z),dz[r],nod[r])));
          float2 C=ffma2(bc(s.x),m2ox[r],ffma2(bc(s.y),m2oy[r],ffma2(bc(s.z),m2oz[r],fadd2(bc(s.w),oo[r]))));
          float2 disc=ffma2(hb,hb,neg2(C));
          m0[r]=__funnelshift_l(__float_as_uint(disc.x), m0[r], 1);
          m1[r]=__funnelshift_l(__float_as_uint(disc.y), m1[r], 1);
        }
      }
      #pragma unroll
      for(int r=0;r<RP;r++){ acc+=__popc(~m0[r])+__popc(~m1[r]); oo[r].x+=1e-6f; }
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
static float time_ms(cudaEvent_t a, cudaEvent_t b){ float ms; CK(cudaEventElapsedTime(&ms,a,b)); return ms; }

int main(int argc, char** argv){
  int dev=0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int sms=p.multiProcessorCount; int clk=0; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,dev);
  printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n",p.name,sms,clk);
  const int threads=256; const int ctas=sms*8;
  float* out; CK(cudaMalloc(&out,sizeof(float)*threads*ctas));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int n=544;
  std::vector<float> h(4*n);
  for(int i=0;i<n;i++){ h[i]=(i%23)-11+0.3f; h[n+i]=0.2f; h[2*n+i]=(i/23)-11+0.4f; h[3*n+i]=0.04f; }
  float* g; CK(cudaMalloc(&g,sizeof(float)*4*n)); CK(cudaMemcpy(g,h.data(),sizeof(float)*4*n,cudaMemcpyHostToDevice));
  { std::vector<float4> c(CN); for(int i=0;i<CN;i++){ int j=i%n; c[i]=make_float4(h[j],h[n+j],h[2*n+j],h[3*n+j]); } CK(cudaMemcpyToSymbol(c_sph,c.data(),sizeof(float4)*CN)); }

#define RUN(name, launch, work_per_thread_flops, extra) do{ \
    for(int rep=0;rep<2;rep++){ launch; } CK(cudaDeviceSynchronize()); \
    float best=1e30f; for(int rep=0;rep<5;rep++){ CK(cudaEventRecord(e0)); launch; CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms=time_ms(e0,e1); if(ms<best)best=ms; } \
    CK(cudaGetLastError()); \
    double fl=(double)(work_per_thread_flops)*threads*ctas; \
    printf("{\"probe\":\"%s\",\"ms\":%.4f,\"tflops\":%.3f%s}\n",name,best,fl/best*1e-9,extra); fflush(stdout); }while(0)

  int iters=4096;
  RUN("ffma_scalar", (k_ffma<<<ctas,threads>>>(out,iters,1.0001f,0.5f)), 2.0*64*iters, "");
  RUN("ffma2_packed", (k_ffma2<<<ctas,threads>>>(out,iters,1.0001f,0.5f)), 4.0*64*iters, "");
  RUN("ffma2_form0_pair_pair_pair", (k_ffma2_forms<0><<<ctas,threads>>>(out,iters,1.0001f,0.5f)), 4.0*64*iters, "");
  RUN("ffma2_form1_pair_scalar_pair", (k_ffma2_forms<1><<<ctas,threads>>>(out,iters,1.0001f,0.5f)), 4.0*64*iters, "");
  RUN("ffma2_plus_shf", (k_ffma2_shf<<<ctas,threads>>>(out,iters,1.0001f,0.5f)), 4.0*64*iters, ",\"note\":\"1 SHF per 2 FFMA2\"");
  { int it=64; 
    for(int rep=0;rep<2;rep++) k_lds128<<<ctas,threads>>>(out,it); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_lds128<<<ctas,threads>>>(out,it); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms=time_ms(e0,e1);
    double lds=(double)1024*it*(threads/32)*ctas; // warp-level LDS.128 instructions
    printf("{\"probe\":\"lds128_broadcast\",\"ms\":%.4f,\"warp_lds_per_clk_per_sm_at_max_clock\":%.4f}\n",ms,lds/(ms*1e-3)/( (double)clk*1e3)/sms); }
  int sit=64;
  size_t smem=sizeof(float)*4*n;
#define SCAN(name,launch,R) RUN(name,launch,17.0*n*sit*(R),",\"flop_per_test\":17")
  SCAN("scan_smem_packed_R1",(k_scan_smem<1><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan_smem_packed_R2",(k_scan_smem<2><<<ctas,threads,smem>>>(g,out,n,sit)),2);
  SCAN("scan_smem_packed_R4",(k_scan_smem<4><<<ctas,threads,smem>>>(g,out,n,sit)),4);
  SCAN("scan_smem8_packed_R1",(k_scan_smem8<1><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan_smem8_packed_R2",(k_scan_smem8<2><<<ctas,threads,smem>>>(g,out,n,sit)),2);
  SCAN("scan_smem8_packed_R4",(k_scan_smem8<4><<<ctas,threads,smem>>>(g,out,n,sit)),4);
  SCAN("scan8_abl_full",(k_scan8_abl<0><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_abl_noshf",(k_scan8_abl<1><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_abl_nolds",(k_scan8_abl<2><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_abl_neither",(k_scan8_abl<3><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_regs_nolds",(k_scan8_regs<<<ctas,threads>>>(g,out,n,sit)),1);
  SCAN("scan8_opmajor_2",(k_scan8_opmajor<2><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_opmajor_4",(k_scan8_opmajor<4><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_opmajor_8",(k_scan8_opmajor<8><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_const_RP1",(k_scan8_const<1><<<ctas,threads>>>(out,n,sit)),2);
  SCAN("scan8_const_RP2",(k_scan8_const<2><<<ctas,threads>>>(out,n,sit)),4);
  SCAN("scan8_constur_RP1",(k_scan8_const_ur<1><<<ctas,threads>>>(out,n,sit)),2);
  SCAN("scan8_constur_RP3",(k_scan8_const_ur<3><<<ctas,threads>>>(out,n,sit)),6);
  SCAN("scan8_pf1",(k_scan_smem8_pf<1><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_pf2",(k_scan_smem8_pf<2><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan8_pf3",(k_scan_smem8_pf<3><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  // occupancy sweep of the 8-op scan: pad dynamic shared memory so that only `occ` CTAs (8 warps each) fit per SM
  for (int occ = 8; occ >= 8; --occ) {
    size_t pad = (size_t)(227 * 1024) / occ - 1024; if (pad < smem) pad = smem; if (pad > 200 * 1024) pad = 200 * 1024;
    CK(cudaFuncSetAttribute(k_scan_smem8<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
    CK(cudaFuncSetAttribute(k_scan_smem8<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
    char nm[64]; snprintf(nm, sizeof nm, "scan8_R1_occ%d", occ);
    SCAN(nm,(k_scan_smem8<1><<<ctas,threads,pad>>>(g,out,n,sit)),1);
    snprintf(nm, sizeof nm, "scan8_R2_occ%d", occ);
    SCAN(nm,(k_scan_smem8<2><<<ctas,threads,pad>>>(g,out,n,sit)),2);
  }
  SCAN("scan_smem_scalar_R1",(k_scan_smem_scalar<1><<<ctas,threads,smem>>>(g,out,n,sit)),1);
  SCAN("scan_smem_scalar_R2",(k_scan_smem_scalar<2><<<ctas,threads,smem>>>(g,out,n,sit)),2);
  SCAN("scan_smem_scalar_R4",(k_scan_smem_scalar<4><<<ctas,threads,smem>>>(g,out,n,sit)),4);
  SCAN("scan_const_packed_RP1",(k_scan_const<1><<<ctas,threads>>>(out,n,sit)),2);
  SCAN("scan_const_packed_RP2",(k_scan_const<2><<<ctas,threads>>>(out,n,sit)),4);
  return 0;
}
