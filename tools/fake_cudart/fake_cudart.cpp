// DIAGNOSTIC TOOL (not product, not oracle): a fake libcudart.so.12 for running the REFERENCE's native prover
// (oracle/_ref/libzprize_ref*.so) on a machine without a GPU, to look at its HOST-side memory behaviour.
//   * "device" memory is host memory that is never reused: every copy / memset is checked against the table of live
//     allocations, so a use-after-free or an overrun of a device buffer is reported with a backtrace;
//   * pinned host buffers carry a 4 KiB guard zone that is checked on cudaFreeHost;
//   * kernels are no-ops (the proof bytes are meaningless);
//   * SHIM_NODE=<bytes>: cudaFree allocates a small bookkeeping record that stays alive, the way a driver keeps free-list
//     nodes.  With 56 bytes (the malloc size class of a make_shared<SyncedMemorykernel> control block) the record lands
//     in the control block the reference has just released, and the reference's SECOND destruction of the same
//     shared_ptr (quotient.cu:277-278,323) then corrupts it: the unmodified library dies at the quotient step,
//     the patched one (oracle/build_pnp_ref.sh) runs clean.  See tools/fake_cudart/run.sh and
//     profiles/r02u_pnp_reference_fake_cudart.log.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <execinfo.h>
#include <cuda_runtime_api.h>
static std::map<char*, std::pair<size_t,bool>> g; // base -> (size, live)
static void bt(){ void* a[32]; int n=backtrace(a,32); backtrace_symbols_fd(a,n,2); }
static int klass(const void* p, size_t n){ // 2 live device, 1 freed device, 0 unknown(host)
  auto it=g.upper_bound((char*)p); if(it==g.begin()) return 0; --it;
  char* b=it->first; size_t s=it->second.first;
  if((char*)p>=b && (char*)p<b+s){ if(!it->second.second) return 1; if((char*)p+n>b+s){fprintf(stderr,"[shim] OVERRUN %p+%zu beyond alloc %p+%zu\n",p,n,b,s); bt(); } return 2; }
  return 0; }
static void check(const char* what,const void* p,size_t n){ if(klass(p,n)==1){ fprintf(stderr,"[shim] USE-AFTER-FREE in %s: %p (%zu bytes)\n",what,p,n); bt(); fflush(stderr); abort(); } }
extern "C" {
cudaError_t cudaMalloc(void** p,size_t n){ char* q=(char*)malloc(n?n:1); memset(q,0x5a,n); g[q]={n,true}; *p=q; return cudaSuccess; }
cudaError_t cudaFree(void* p){ if(!p) return cudaSuccess; auto it=g.find((char*)p); if(it==g.end()||!it->second.second){fprintf(stderr,"[shim] BAD FREE %p\n",p); bt(); abort();} it->second.second=false; memset(p,0xdd,it->second.first); if(getenv("SHIM_NODE")){ long* node=(long*)malloc(atoi(getenv("SHIM_NODE"))); node[0]=(long)p; node[1]=1; } return cudaSuccess; } // never really freed: addresses stay unique
static std::map<void*,size_t> hs;
cudaError_t cudaMallocHost(void** p,size_t n){ char* q=(char*)malloc(n+4096); memset(q+n,0xa5,4096); hs[q]=n; *p=q; return cudaSuccess; }
cudaError_t cudaFreeHost(void* p){ auto it=hs.find(p); if(it!=hs.end()){ size_t n=it->second; unsigned char* q=(unsigned char*)p; size_t last=0; for(size_t i=0;i<4096;i++) if(q[n+i]!=0xa5) last=i+1; if(last){ fprintf(stderr,"[shim] PINNED HOST OVERRUN: %zu-byte buffer written %zu bytes past its end\n",n,last); bt(); } hs.erase(it);} free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d,const void* s,size_t n,cudaMemcpyKind){ check("cudaMemcpy src",s,n); check("cudaMemcpy dst",d,n); memmove(d,s,n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d,const void* s,size_t n,cudaMemcpyKind k,cudaStream_t){ return cudaMemcpy(d,s,n,k); }
cudaError_t cudaMemcpy2D(void* d,size_t dp,const void* s,size_t sp,size_t w,size_t h,cudaMemcpyKind){ for(size_t i=0;i<h;i++){ check("cudaMemcpy2D src",(char*)s+i*sp,w); check("cudaMemcpy2D dst",(char*)d+i*dp,w); memmove((char*)d+i*dp,(char*)s+i*sp,w);} return cudaSuccess; }
cudaError_t cudaMemset(void* d,int v,size_t n){ check("cudaMemset",d,n); memset(d,v,n); return cudaSuccess; }
cudaError_t cudaGetDevice(int* d){ *d=0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p,int){ memset(p,0,sizeof(*p)); p->multiProcessorCount=148; p->major=10; p->maxThreadsPerBlock=1024; p->sharedMemPerBlock=49152; p->sharedMemPerBlockOptin=232448; p->warpSize=32; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v,cudaDeviceAttr a,int){ *v = (a==cudaDevAttrMultiProcessorCount)?148:(a==cudaDevAttrMaxSharedMemoryPerBlockOptin?232448:1024); return cudaSuccess; }
cudaError_t cudaGetLastError(){ return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t){ return "fake"; }
cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a,const void*){ memset(a,0,sizeof(*a)); a->maxThreadsPerBlock=1024; a->numRegs=64; return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n,const void*,int,size_t,unsigned){ *n=2; return cudaSuccess; }
cudaError_t cudaLaunchKernel(const void*,dim3,dim3,void**,size_t,cudaStream_t){ return cudaSuccess; }
cudaError_t cudaLaunchCooperativeKernel(const void*,dim3,dim3,void**,size_t,cudaStream_t){ return cudaSuccess; }
unsigned __cudaPushCallConfiguration(dim3,dim3,size_t,void*){ return 0; }
cudaError_t __cudaPopCallConfiguration(dim3*,dim3*,size_t*,void*){ return cudaSuccess; }
void** __cudaRegisterFatBinary(void*){ static void* h; return &h; }
void __cudaRegisterFatBinaryEnd(void**){}
void __cudaUnregisterFatBinary(void**){}
void __cudaRegisterFunction(void**,const char*,char*,const char*,int,uint3*,uint3*,dim3*,dim3*,int*){}
void __cudaRegisterVar(void**,char*,char*,const char*,int,size_t,int,int){}
}
