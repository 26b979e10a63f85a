"""jxl_coder_b200 — B200-native drop-in for the JPEG XL decode path of awxkee/jxl-coder.

Host-side mirror of the reference's Kotlin surface (JxlCoder.decode / decodeSampled / getSize, JxlAnimatedImage;
/root/reference/jxlcoder/src/main/java/com/awxkee/jxlcoder/JxlCoder.kt:50-105,191-193 and JxlAnimatedImage.kt:43-193)
over the C ABI of libjxlb200.so (include/jxlb200.h).  All pixels come from CUDA kernels; importing this package on a
machine without the built library, or decoding without a CUDA device, raises — there is no CPU path.
"""
from .api import (Bitmap, InvalidJXLException, InvalidImageSizeException, JxlAnimatedImage, JxlCoder, JxlCoderError,
                  JxlResizeFilter, PreferredColorConfig, ScaleMode, UnsupportedJXLException, decode_batch, kernel_launches,
                  last_batch_timings, lib_path, load_library, PendingBatch, PreparedBatch)

__all__ = ["Bitmap", "InvalidJXLException", "InvalidImageSizeException", "JxlAnimatedImage", "JxlCoder", "JxlCoderError",
           "JxlResizeFilter", "PreferredColorConfig", "ScaleMode", "UnsupportedJXLException", "decode_batch", "kernel_launches",
           "last_batch_timings", "lib_path", "load_library", "PendingBatch", "PreparedBatch"]
