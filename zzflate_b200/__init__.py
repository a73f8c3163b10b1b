"""zzflate_b200 -- B200-native deflate encoder behind the zzflate API.  See DESIGN.md."""
from .api import (Config, Format, ZzFlateDecode, ZzFlateEncode, ZzFlateEncodeToCallback, adler32x, bound, combine, crc32,
                  crc32_combine, debug_chunk, deflate_device, deflate_raw, device_count, encode_ptr)
from ._lib import Stats, ZzGpuError

__all__ = ["Config", "Format", "ZzFlateDecode", "ZzFlateEncode", "ZzFlateEncodeToCallback", "adler32x", "bound", "combine", "crc32",
           "crc32_combine", "debug_chunk", "deflate_device", "deflate_raw", "device_count", "encode_ptr", "Stats", "ZzGpuError"]
