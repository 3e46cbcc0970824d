// FP64 pipe micro-benchmarks for B200 (sm_100a): DMMA m8n8k4 / m16n8k16 issue rate vs DFMA, per SM and full chip.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int NACC>
__global__ void __launch_bounds__(256) dmma884(double* out, int iters, double a0, double b0){
    double a=a0+threadIdx.x*1e-9, b=b0;
    double c[NACC][2];
    #pragma unroll
    for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
    for(int it=0; it<iters; ++it){
        #pragma unroll
        for(int i=0;i<NACC;i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(256) dmma16816(double* out, int iters, double a0, double b0){
    double a[8], b[4];
    #pragma unroll
    for(int i=0;i<8;i++) a[i]=a0+i+threadIdx.x*1e-9;
    #pragma unroll
    for(int i=0;i<4;i++) b[i]=b0+i;
    double c[NACC][4];
    #pragma unroll
    for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;c[i][2]=0;c[i][3]=0;}
    for(int it=0; it<iters; ++it){
        #pragma unroll
        for(int i=0;i<NACC;i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(256) dfma(double* out, int iters, double a0, double b0){
    double a=a0+threadIdx.x*1e-9, b=b0;
    double c[NACC];
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=i;
    for(int it=0; it<iters; ++it){
        #pragma unroll
        for(int i=0;i<NACC;i++) c[i]=fma(a,c[i],b);
    }
    double s=0;
    #pragma unroll
    for(int i=0;i<NACC;i++) s+=c[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void dexp(double* out, int iters, double x0){
    double x=x0+threadIdx.x*1e-3, s=0;
    for(int it=0; it<iters; ++it){ s+=exp(-x); x+=1e-6; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<class F> float timeit(F f){
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    printf("device %s SMs %d clock %d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
    double* out; CK(cudaMalloc(&out, 148*32*1024*8));
    int iters=4096;
    for(int blocks_per_sm: {1,2,4}) for(int threads: {128,256}){
        int grid=p.multiProcessorCount*blocks_per_sm;
        { float ms=timeit([&]{dmma884<16><<<grid,threads>>>(out,iters,1.0,1e-3);});
          double fl=(double)grid*(threads/32)*iters*16*512.0; printf("DMMA884  acc16 bps=%d thr=%d: %.3f ms %.2f TFLOP/s\n",blocks_per_sm,threads,ms,fl/ms*1e-9);}
        { float ms=timeit([&]{dmma884<32><<<grid,threads>>>(out,iters,1.0,1e-3);});
          double fl=(double)grid*(threads/32)*iters*32*512.0; printf("DMMA884  acc32 bps=%d thr=%d: %.3f ms %.2f TFLOP/s\n",blocks_per_sm,threads,ms,fl/ms*1e-9);}
        { float ms=timeit([&]{dmma16816<8><<<grid,threads>>>(out,iters,1.0,1e-3);});
          double fl=(double)grid*(threads/32)*iters*8*4096.0; printf("DMMA16816 acc8 bps=%d thr=%d: %.3f ms %.2f TFLOP/s\n",blocks_per_sm,threads,ms,fl/ms*1e-9);}
        { float ms=timeit([&]{dfma<16><<<grid,threads>>>(out,iters*8,1.0000001,1e-9);});
          double fl=(double)grid*threads*iters*8*16*2.0; printf("DFMA     acc16 bps=%d thr=%d: %.3f ms %.2f TFLOP/s\n",blocks_per_sm,threads,ms,fl/ms*1e-9);}
    }
    { int grid=148*8, threads=256; float ms=timeit([&]{dexp<<<grid,threads>>>(out,4096,0.5);});
      printf("exp(double): %.3f ms %.2f Gexp/s\n",ms,(double)grid*threads*4096/ms*1e-6);}
    CK(cudaDeviceSynchronize());
    return 0;
}
