// Developer probe: back-to-back dependent kernels on one stream, ordinary launches against programmatic dependent launch
// (griddepcontrol.wait at the top of the kernel).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a pdl_probe.cu -o pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void Work(float* a, int n, int rounds)
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		float v = a[i];
		for (int k = 0; k < rounds; ++k) v = v * 1.0001f + 0.5f;
		a[i] = v;
	}
}

int main()
{
	const int n = 1184 * 256;
	float* a;
	cudaMalloc(&a, n * sizeof(float));
	cudaMemset(a, 0, n * sizeof(float));
	cudaStream_t s;
	cudaStreamCreate(&s);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int rounds : {8, 200, 2000})
		for (int mode = 0; mode < 2; ++mode)
		{
			float best = 1e9f;
			for (int rep = 0; rep < 5; ++rep)
			{
				cudaEventRecord(e0, s);
				for (int k = 0; k < 200; ++k)
				{
					if (mode == 0) Work<<<1184, 256, 0, s>>>(a, n, rounds);
					else
					{
						cudaLaunchConfig_t cfg = {};
						cfg.gridDim = dim3(1184);
						cfg.blockDim = dim3(256);
						cfg.stream = s;
						cudaLaunchAttribute at[1];
						at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
						at[0].val.programmaticStreamSerializationAllowed = 1;
						cfg.attrs = at;
						cfg.numAttrs = 1;
						cudaLaunchKernelEx(&cfg, Work, a, n, rounds);
					}
				}
				cudaEventRecord(e1, s);
				cudaEventSynchronize(e1);
				float ms;
				cudaEventElapsedTime(&ms, e0, e1);
				if (ms < best) best = ms;
			}
			printf("rounds %d %s: %.2f us per launch\n", rounds, mode ? "PDL" : "plain", 1e3f * best / 200);
		}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
