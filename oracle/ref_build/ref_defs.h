/* Force-included into every reference translation unit: the two string macros
 * the reference normally gets from simulation/CMakeLists.txt:30-31. */
#pragma once
#ifndef PROJECT_DIR
#define PROJECT_DIR "/root/reference/simulation"
#endif
#ifndef GRAPHICS_CODEBASE_SOURCE_DIR
#define GRAPHICS_CODEBASE_SOURCE_DIR "/root/reference/simulation"
#endif
