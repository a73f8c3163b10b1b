/* huffman.h -- host-side canonical-code helpers with the reference's names (zzflate/huffman.h:44-81,
 * huffman.cpp:11-43): huffman::reverse, huffman::defaultTableLengths, huffman::generate<T>.  The GPU builds the
 * per-chunk codes in K-HUFF; these host versions serve callers and the reference's known-answer tests
 * (zztest/TestHuffman.cpp:34-50, TestBitOutput.cpp:40-48).  New code written from RFC 1951 3.2.2 / SURVEY A.5. */
#ifndef ZZFLATE_B200_HUFFMAN_H
#define ZZFLATE_B200_HUFFMAN_H

#include <stdint.h>
#include <vector>

namespace huffman
{
const int MAX_BITS = 16;                                   /* huffman.h:44 */

/* the low `length` bits of `value`, mirrored (codes are stored ready for an LSB-first bit writer) */
inline unsigned reverse(unsigned value, int length)
{
    unsigned r = 0;
    for (int i = 0; i < length; ++i) r |= ((value >> i) & 1u) << (length - 1 - i);
    return r;
}

/* code lengths of the fixed Huffman table, RFC 1951 3.2.6: 8 x144, 9 x112, 7 x24, 8 x8 */
inline std::vector<int> defaultTableLengths()
{
    std::vector<int> l(288);
    for (int i = 0; i < 288; ++i) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
    return l;
}

/* canonical codes from lengths (RFC 1951 3.2.2), stored bit-reversed; entries of length 0 are left untouched */
template <class T>
void generate(const std::vector<int>& lengths, T* codes)
{
    int count[MAX_BITS] = {};
    for (int len : lengths) count[len]++;
    count[0] = 0;
    unsigned next[MAX_BITS] = {};
    unsigned c = 0;
    for (int bits = 1; bits < MAX_BITS; ++bits) { c = (c + (unsigned)count[bits - 1]) << 1; next[bits] = c; }
    for (size_t n = 0; n < lengths.size(); ++n) {
        const int len = lengths[n];
        if (len == 0) continue;
        codes[n] = T{ len, reverse(next[len]++, len) };
    }
}
}  // namespace huffman

#endif
