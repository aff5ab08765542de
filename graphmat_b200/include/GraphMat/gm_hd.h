// gm_hd.h -- the one annotation a vertex program needs to run on the device.
//
// The reference calls the four operators through virtual dispatch on a host
// object (include/GraphProgram.h:73-99 of narayanan2004/GraphMat).  The device
// engine calls them by qualified name on a by-value copy of the program, so the
// bodies stay byte-identical to the reference's; only GM_HD is added in front.
#ifndef GM_HD_H
#define GM_HD_H
#if defined(__CUDACC__)
#define GM_HD __host__ __device__
#else
#define GM_HD
#endif
#endif
