// Float formatting of the statistics lines (fastaToKmerCoverageStats.cpp:140-148 prints through an ostream with default
// settings: precision 6, %g).
#pragma once
#include <math.h>
#include <stdio.h>

#include <charconv>

// iostream default float formatting (precision 6, %g); x86 default NaN carries the sign bit -> "-nan"
// (std::to_chars(general, 6) is specified to print what printf("%g") prints in the C locale, at less than half its cost --
// and two of these per read are the most expensive thing left on the host; checked against sprintf on 2e7 floats of every
// kind, tests/test_cabi_and_host.py keeps a sample)
inline int fmt_float(char* out, float f) {
    if (isnan(f)) return sprintf(out, signbit(f) ? "-nan" : "nan");
    if (isinf(f)) return sprintf(out, "%g", (double)f);
    return (int)(std::to_chars(out, out + 32, (double)f, std::chars_format::general, 6).ptr - out);
}

