#include <cuda_runtime.h>
extern "C" __global__ void k_vimax2(unsigned *o, unsigned a, unsigned b){ o[0] = __vimax3_s16x2(a,b,b); }
extern "C" __global__ void k_viaddmax2(unsigned *o, unsigned a, unsigned b, unsigned c){ o[0] = __viaddmax_s16x2(a,b,c); }
extern "C" __global__ void k_vibmax2(unsigned *o, unsigned a, unsigned b){ bool h,l; o[0] = __vibmax_s16x2(a,b,&h,&l); o[1]=h; o[2]=l; }
extern "C" __global__ void k_vadd2(unsigned *o, unsigned a, unsigned b){ o[0] = __vadd2(a,b); }
extern "C" __global__ void k_vsub2(unsigned *o, unsigned a, unsigned b){ o[0] = __vsub2(a,b); }
extern "C" __global__ void k_vcmpgts2(unsigned *o, unsigned a, unsigned b){ o[0] = __vcmpgts2(a,b); }
extern "C" __global__ void k_vmaxs4(unsigned *o, unsigned a, unsigned b){ o[0] = __vmaxs4(a,b); }
extern "C" __global__ void k_vmaxu4(unsigned *o, unsigned a, unsigned b){ o[0] = __vmaxu4(a,b); }
extern "C" __global__ void k_vadd4(unsigned *o, unsigned a, unsigned b){ o[0] = __vadd4(a,b); }
extern "C" __global__ void k_vsub4(unsigned *o, unsigned a, unsigned b){ o[0] = __vsub4(a,b); }
extern "C" __global__ void k_vcmpgtu4(unsigned *o, unsigned a, unsigned b){ o[0] = __vcmpgtu4(a,b); }
extern "C" __global__ void k_vcmpgts4(unsigned *o, unsigned a, unsigned b){ o[0] = __vcmpgts4(a,b); }
extern "C" __global__ void k_vmaxs2(unsigned *o, unsigned a, unsigned b){ o[0] = __vmaxs2(a,b); }
extern "C" __global__ void k_vmaxu2(unsigned *o, unsigned a, unsigned b){ o[0] = __vmaxu2(a,b); }
extern "C" __global__ void k_vimax3_2(unsigned *o, unsigned a, unsigned b, unsigned c){ o[0] = __vimax3_s16x2(a,b,c); }
extern "C" __global__ void k_vimin2u(unsigned *o, unsigned a, unsigned b){ o[0] = __vimin3_u16x2(a,b,b); }
