/* rmq.cu -- placeholder: flags reads that need the long-join re-chain. */
#include "dev_common.cuh"
#include "stages.h"

__global__ void rechain_kernel(ChunkDev c, DevOpt o, uint32_t r0, uint32_t r1, uint32_t *work)
{
	for (uint32_t r = r0 + blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += gridDim.x * blockDim.x) {
		const int n_u = (int)c.n_u[r];
		if (o.bw_long > o.bw && !(o.flag & MMG_F_NO_LJOIN) && n_u > 1) {
			const uint64_t ab = c.a_off[r] - c.a_off0;
			const int qlen = (int)(c.off[r + 1] - c.off[r]);
			int32_t st = (int32_t)c.by[ab], en = (int32_t)c.by[ab + (uint32_t)c.u[ab] - 1];
			if (qlen - (en - st) > o.rmq_rescue_size || (float)(en - st) > __fmul_rn((float)qlen, o.rmq_rescue_ratio))
				c.flags[r] |= 2u;
		}
	}
}

int launch_rechain(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work)
{
	MMG_LAUNCH(rechain_kernel, n_sms, 128, 0, st, c, o, r0, r1, work);
	return 0;
}
