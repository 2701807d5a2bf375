// wire.hpp -- output wire formats of the reference CLI's `pipe` subcommand (src/main.rs:668-778, 1106-1133,
// 1268-1286), restated for batched results.  Host-only, header-only; used by yume_pdq_pipe.cpp.
//
// A record is one hash with its quality (and, for `diagnostic`, threshold and PDQF):
//   raw            32 hash bytes, byte 0 first (main.rs:684-686)
//   q+raw          f32 little-endian quality, then the 32 bytes (main.rs:687-690)
//   hex / HEX      64 hex digits, byte 0 first, high nibble first (main.rs:691-720)
//   bin / BIN      256 characters '0'/'1', byte 0 first, most significant bit first (main.rs:721-737,
//                  BINARY_PRINTING = "{:08b}", build.rs:504-508)
//   q+hex q+HEX q+bin q+BIN   the same behind "{:02.3}:" of quality * 100 (f32 product) (main.rs:693, 723)
//   diagnostic     multi-line dump (main.rs:738-774)
// followed by the separator selected by a trailing +lf / +crlf / +cr / +nul / +null / +nl (main.rs:1106-1133).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace ypdq_wire {

enum OutputType : int {
    RAW = 0, RAW_PREFIX_QUALITY = 1, ASCII_HEX = 2, ASCII_BINARY = 3, ASCII_HEX_PREFIX_QUALITY = 4,
    ASCII_BINARY_PREFIX_QUALITY = 5, DIAGNOSTIC = 6, FLAG_UPPER = 128,    // main.rs:348-358
};

struct Format {
    int type = ASCII_BINARY;
    std::string separator;
};

// main.rs:1106-1133 (separator = last '+' component, if it names one) and main.rs:1268-1286 (the rest, exact match)
inline bool parse_format(const std::string &spec, Format *out)
{
    std::string body = spec, sep;
    const size_t plus = spec.rfind('+');
    const std::string last = plus == std::string::npos ? spec : spec.substr(plus + 1);
    bool has_sep = true;
    if (last == "nl" || last == "NL" || last == "lf" || last == "LF")
        sep = "\n";                      // "nl" = the platform's newline; this build targets Linux
    else if (last == "crlf" || last == "CRLF")
        sep = "\r\n";
    else if (last == "cr" || last == "CR")
        sep = "\r";
    else if (last == "nul" || last == "NUL" || last == "null" || last == "NULL")
        sep = std::string(1, '\0');
    else
        has_sep = false;
    if (has_sep)
        body = plus == std::string::npos ? std::string() : spec.substr(0, plus);
    int type;
    if (body == "raw" || body == "Raw" || body == "RAW")
        type = RAW;
    else if (body == "q+raw" || body == "q+Raw" || body == "q+RAW")
        type = RAW_PREFIX_QUALITY;
    else if (body == "hex")
        type = ASCII_HEX;
    else if (body == "HEX")
        type = ASCII_HEX | FLAG_UPPER;
    else if (body == "bin" || body == "BIN")
        type = ASCII_BINARY;
    else if (body == "q+hex")
        type = ASCII_HEX_PREFIX_QUALITY;
    else if (body == "q+HEX")
        type = ASCII_HEX_PREFIX_QUALITY | FLAG_UPPER;
    else if (body == "q+bin" || body == "q+BIN")
        type = ASCII_BINARY_PREFIX_QUALITY;
    else if (body == "diagnostic")
        type = DIAGNOSTIC;
    else
        return false;
    out->type = type;
    out->separator = sep;
    return true;
}

// Rust's `{:.N}` of an f32: exact decimal expansion of the value, correctly rounded; glibc's printf of the same
// value widened to double prints the same digits.  Rust spells NaN "NaN" and never prints a sign for it.
inline void append_fixed(std::string &s, float v, int decimals)
{
    if (std::isnan(v)) {
        s += "NaN";
        return;
    }
    char buf[96];
    const int n = snprintf(buf, sizeof buf, "%.*f", decimals, (double)v);
    s.append(buf, (size_t)n);
}

inline void append_quality_prefix(std::string &s, float quality)
{
    append_fixed(s, quality * 100.0f, 3);    // write!("{0:02.3}:", quality * 100.0): the width never binds
    s += ':';
}

inline void append_bin(std::string &s, const uint8_t *hash)
{
    for (int i = 0; i < 32; i++)
        for (int b = 7; b >= 0; b--)
            s += (char)('0' + ((hash[i] >> b) & 1));
}

struct Record {
    const uint8_t *hash;     // 32 bytes
    float quality;
    float threshold;         // diagnostic: the threshold this record was quantised with
    const float *pdqf;       // diagnostic: 256 f32, the (transformed) PDQF of this record
    float star_threshold;    // diagnostic: entries above it get a '*' -- the ORIGINAL threshold for all 8 records
                             // (macro hygiene: the macro body's `threshold` is the outer binding, main.rs:766)
    uint32_t dihedral;       // packed Dihedrals value; 0 = the "original" record
    unsigned index;          // diagnostic: 0 for the original, 1..7 for the dihedrals of the same frame
};

inline void append_record(std::string &s, const Format &f, const Record &r)
{
    const int type = f.type & ~FLAG_UPPER;
    switch (type) {
    case RAW_PREFIX_QUALITY: {
        uint32_t bits;
        memcpy(&bits, &r.quality, 4);
        for (int i = 0; i < 4; i++)
            s += (char)((bits >> (8 * i)) & 255);   // f32::to_le_bytes
    }
        [[fallthrough]];
    case RAW:
        s.append(reinterpret_cast<const char *>(r.hash), 32);
        break;
    case ASCII_HEX_PREFIX_QUALITY:
        append_quality_prefix(s, r.quality);
        [[fallthrough]];
    case ASCII_HEX: {
        const char *digits = (f.type & FLAG_UPPER) ? "0123456789ABCDEF" : "0123456789abcdef";
        for (int i = 0; i < 32; i++) {
            s += digits[r.hash[i] >> 4];
            s += digits[r.hash[i] & 15];
        }
        break;
    }
    case ASCII_BINARY_PREFIX_QUALITY:
        append_quality_prefix(s, r.quality);
        [[fallthrough]];
    case ASCII_BINARY:
        append_bin(s, r.hash);
        break;
    case DIAGNOSTIC: {
        if (r.dihedral == 0)
            s += "[\"original\"]\n";               // {:?} of a &str
        else
            s += "[Dihedrals { packed: " + std::to_string(r.dihedral) + " }]\n";   // #[derive(Debug)], lib.rs:85-93
        s += "Index: " + std::to_string(r.index) + "\n";
        s += "Quality: ";
        append_fixed(s, r.quality * 100.0f, 3);
        s += "\nThreshold: ";
        append_fixed(s, r.threshold, 7);
        s += "\nHash:\n";
        append_bin(s, r.hash);
        s += "\nPDQF:\n";
        for (int i = 0; i < 16; i++) {
            for (int j = 0; j < 16; j++) {
                if (j > 0)
                    s += '\t';
                const float v = r.pdqf[i * 16 + j];
                if (v > r.star_threshold)
                    s += '*';
                append_fixed(s, v, 5);
            }
            s += '\n';
        }
        break;
    }
    }
    s += f.separator;
}

// PDQF state of dihedral d (1..7, YUME_PDQ_CUDA_DIHEDRAL_ORDER) from the untransformed PDQF: the in-place
// sequence of lib.rs:391-417 leaves C.P, R.C.P, R.P, P^T, R.P^T, C.R.P^T, C.P^T (C / R = negate even columns / rows)
inline void dihedral_pdqf(const float *pdqf, int d, float *out)
{
    const bool t = d & 4, c = (0xC6 >> d) & 1, r = (0x6C >> d) & 1;
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j++) {
            const float v = t ? pdqf[j * 16 + i] : pdqf[i * 16 + j];
            const bool neg = (c && !(j & 1)) != (r && !(i & 1));
            out[i * 16 + j] = neg ? -v : v;
        }
}

}  // namespace ypdq_wire
