// Tensor maps (TMA descriptors) for the pitch-padded field buffers of the fast 3-D path.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Encode a tiled tensor map over complex<float> data viewed as float32:
//   dims[0..rank)   extents, innermost first, dims[0] in COMPLEX elements (doubled internally)
//   strides[1..rank) byte strides of dims 1.. (multiples of 16)
//   box[0..rank)    box extents, box[0] in complex elements
// Returns 0 on success; *err receives a static message otherwise.
int exb_tma_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char** err);
