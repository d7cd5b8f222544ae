/* Base integer / float aliases the correspondence headers use (u8 .. u64, s8 .. s64, f32, f64, usize,
 * u32_max). Same names and meanings as the reference's include/vslam_internal.h:9-32, so code written
 * against it (Frame.h, PointMap.h, vslam.cpp) compiles unchanged against this repo's headers. */
#ifndef VSLAM_B200_VSLAM_INTERNAL_H
#define VSLAM_B200_VSLAM_INTERNAL_H

#include <float.h>
#include <limits.h>
#include <stdint.h>

#include <cstddef>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int8_t s8;
typedef int16_t s16;
typedef int32_t s32;
typedef int64_t s64;
typedef float f32;
typedef double f64;
typedef size_t usize; /* index into memory */

#ifndef u32_max
#define u32_max ((u32)-1)
#endif
#ifndef f32_maximum
#define f32_maximum FLT_MAX
#endif
#ifndef internal_function
#define internal_function static
#define local_persist static
#define global_variable static
#endif

#endif
