// Developer probe: host round trip after a short kernel -- cudaMemcpyAsync(D2H) + cudaStreamSynchronize against a kernel
// that stores the values into page-locked host memory and a host thread that polls a sequence word.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a sync_probe.cu -o sync_probe
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void Work(int* counters, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) atomicAdd(&counters[i & 63], 1);
}
__global__ void Publish(const int* counters, volatile int* host, int seq)
{
	int t = threadIdx.x;
	if (t < 64) host[t] = counters[t];
	__threadfence_system();
	__syncthreads();
	if (t == 0) host[64] = seq;
}

int main()
{
	int* counters;
	cudaMalloc(&counters, 256);
	cudaMemset(counters, 0, 256);
	int* host;
	cudaMallocHost(&host, 512);
	host[64] = 0;
	cudaStream_t s;
	cudaStreamCreate(&s);
	typedef std::chrono::steady_clock Clock;
	for (int mode = 0; mode < 2; ++mode)
	{
		double best = 1e9;
		for (int rep = 0; rep < 5; ++rep)
		{
			cudaStreamSynchronize(s);
			Clock::time_point t0 = Clock::now();
			for (int k = 1; k <= 200; ++k)
			{
				Work<<<148, 256, 0, s>>>(counters, 148 * 256);
				if (mode == 0)
				{
					cudaMemcpyAsync(host, counters, 256, cudaMemcpyDeviceToHost, s);
					cudaStreamSynchronize(s);
				}
				else
				{
					const int seq = rep * 1000 + k;
					Publish<<<1, 64, 0, s>>>(counters, host, seq);
					while (((volatile int*)host)[64] != seq) {}
				}
			}
			double us = std::chrono::duration<double, std::micro>(Clock::now() - t0).count() / 200;
			if (us < best) best = us;
		}
		printf("%s: %.2f us per kernel + read-back\n", mode ? "publish kernel + host poll" : "memcpy + stream sync", best);
	}
	return 0;
}
