/* crc.h -- CRC-32 (poly 0xEDB88320) with zlib-style chaining through startValue, as zzflate/crc.h:7 and
 * crc.cpp:24-33; computed on the GPU (zzgpu_checksums). */
#ifndef ZZFLATE_B200_CRC_H
#define ZZFLATE_B200_CRC_H

#include <stddef.h>
#include <stdint.h>

uint32_t crc32(const uint8_t* buffer, size_t length, uint32_t startValue = 0);

#endif
