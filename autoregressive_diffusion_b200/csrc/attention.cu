// placeholder until the attention kernels land
