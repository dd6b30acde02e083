// Tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link against libcuda).
#pragma once

#include <cuda.h>
#include <stdint.h>

namespace cb2 {

// 3-D fp16 map, no swizzle / interleave; dims and box in elements (innermost first), strides in bytes
int encode_tensor_map_3d(CUtensorMap *tm, const void *base, const uint64_t (&dims)[3], const uint64_t (&strides)[2],
                         const uint32_t (&box)[3]);

// K-major operand image [k8][rows][8 halves]; box = box_rows rows x box_k8 K-groups, landing in shared memory as
// [box_k8][box_rows][16 B].  Coordinates: (0, row / 32, k8)
int encode_weight_map(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t k8, uint32_t box_rows, uint32_t box_k8);

}  // namespace cb2
