// Sequence files as `jellyfish count` reads them -> record batches (each sequence followed by one '\n', the layout
// libtrinity_gpu consumes).  FASTA: a record is every line up to the next line starting with '>'; line breaks inside a
// record do NOT break k-mers (the lines are joined), a trailing '\r' is dropped, text before the first header is
// ignored.  FASTQ (first non-blank byte '@'): 4-line records, the sequence is line 2.  `flush` is called whenever the
// batch exceeds flush_bytes (at a record boundary) and once at the end; it consumes and clears `recs`.
#pragma once
#include <stddef.h>
#include <vector>

#include "fasta_io.hpp"

namespace tgio {

template <class Flush>
inline void parse_sequence_file(const char* p, const char* end, std::vector<char>& recs, size_t flush_bytes, Flush flush) {
    while (p < end && (*p == '\n' || *p == '\r')) p++;
    const bool fastq = p < end && *p == '@';
    if (fastq) {
        while (p < end) {
            const char* nl = find_nl(p, end); p = nl < end ? nl + 1 : end;          // @name
            if (p >= end) break;
            nl = find_nl(p, end);
            const char* e = nl;
            if (e > p && e[-1] == '\r') e--;
            recs.insert(recs.end(), p, e); recs.push_back('\n');                  // sequence
            p = nl < end ? nl + 1 : end;
            nl = find_nl(p, end); p = nl < end ? nl + 1 : end;                    // +
            nl = find_nl(p, end); p = nl < end ? nl + 1 : end;                    // qualities
            if (recs.size() > flush_bytes) flush();
        }
    } else {
        bool open = false;
        while (p < end) {
            const char* nl = find_nl(p, end);
            if (*p == '>') {
                if (open) recs.push_back('\n');
                open = true;
                if (recs.size() > flush_bytes) flush();
            } else if (open) {
                const char* e = nl;
                if (e > p && e[-1] == '\r') e--;
                recs.insert(recs.end(), p, e);
            }
            p = nl < end ? nl + 1 : end;
        }
        if (open) recs.push_back('\n');
    }
    if (!recs.empty()) flush();
}

}  // namespace tgio
