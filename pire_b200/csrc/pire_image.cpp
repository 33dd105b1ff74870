// pire_image.cpp -- see pire_image.hpp.
#include "pire_image.hpp"

#include <cstring>

namespace pire_b200 {

namespace {

// Stream header, pire/scanners/common.h:44-63.
struct StreamHeader {
    uint32_t magic, version, ptr_size, max_word_size, type, hdr_size;
};
static_assert(sizeof(StreamHeader) == 24, "header layout");

// Scanner::Locals on x86-64, pire/scanners/multi.h:315-323.
struct Locals {
    uint32_t states, letters, regexps, pad0;
    uint64_t initial;            // byte offset from the first row (multi.h:564)
    uint32_t final_table_size, pad1;
    uint64_t relocation;         // 1 = Relocatable (multi.h:56), 2 = Nonrelocatable (:72)
    uint64_t shortcutting;       // 0x1000 NoShortcuts (:828), 0x2000+N ExitMasks<N> (:700)
};
static_assert(sizeof(Locals) == 48, "locals layout");

constexpr uint32_t kMagic = 0x45524950u;     // "PIRE", common.h:52
constexpr uint64_t kEnd = ~uint64_t(0);      // accept-list terminator, multi.h:96

size_t RoundUp(size_t v, size_t b) { return (v + b - 1) / b * b; }

} // namespace

std::string ParsePireImage(const void* data, size_t size, Dfa* out)
{
    *out = Dfa();
    const uint8_t* p = static_cast<const uint8_t*>(data);
    if (!p)
        return "null scanner image";
    if (size < sizeof(StreamHeader) + sizeof(Locals) + 8)
        return "scanner image truncated (header)";

    StreamHeader h;
    std::memcpy(&h, p, sizeof(h));
    if (h.magic != kMagic || h.ptr_size != 8 || h.max_word_size != 16)
        return "Serialized regexp incompatible with your system";           // common.h:67-68
    if (h.version != 7 && h.version != 6)
        return "You are trying to used an incompatible version of a serialized regexp";  // :69-70
    if (h.type != 1 || h.hdr_size != sizeof(Locals))
        return "Serialized regexp incompatible with your system";           // :71-76 (not a multi Scanner)
    size_t pos = RoundUp(sizeof(h), 8);

    Locals m;
    std::memcpy(&m, p + pos, sizeof(m));
    pos += RoundUp(sizeof(m), 8);
    if (m.relocation != 1)
        return "Type mismatch while mmapping Pire::Scanner";                 // multi.h:257-258
    uint32_t header_cells;
    if (m.shortcutting == 0x1000)
        header_cells = 2;        // CommonRowHeader only: one size_t = two 4-byte cells (:831-833)
    else if (m.shortcutting == 0x2002)
        header_cells = 18;       // 2 masks x 4 size_t + Flags = 72 bytes (:706-767)
    else
        return "This scanner has different shortcutting type";               // :262-263

    const bool empty = p[pos] != 0;                                          // :567
    pos += 8;

    out->exit_masks = header_cells == 18;
    if (empty) {
        // Scanner::Null(): a one-state automaton that accepts nothing (:339-344).
        out->empty = true;
        out->states = 1;
        out->letters = 1;
        out->regexps = 0;                                                    // RegexpsCount() :139
        out->initial = 0;
        out->class_of.assign(kMaxCharUnaligned, 0);
        out->next.assign(1, 0);
        out->flags.assign(1, 2);
        out->acc_begin.assign(2, 0);
        return std::string();
    }

    if (m.states == 0 || m.letters == 0)
        return "scanner image has no states or letters";
    // A scanner has at most MaxChar letter classes (defs.h:71) and Glue caps the states (multi.h:1100); an
    // image that claims more is corrupt.  The bounds also keep the size arithmetic below inside 64 bits, so a
    // crafted header cannot wrap BufSize() around the EOF check.
    if (m.letters > kMaxChar)
        return "scanner image: more letter classes than MaxChar";
    const size_t row_cells = RoundUp((size_t) m.letters + header_cells, 4);   // RowSize() :347
    const size_t row_bytes = row_cells * 4;
    if ((size_t) m.states > size / row_bytes || (size_t) m.final_table_size > size / 8)
        return "EOF reached while mapping Pire::Scanner";                    // :271-272 (tables larger than the image)
    const size_t buf = RoundUp((size_t) kMaxChar * 2 + (size_t) m.final_table_size * 8
                               + (size_t) m.states * 8 + row_bytes * m.states, 8);   // BufSize() :297-305
    if (size < pos || size - pos < buf)
        return "EOF reached while mapping Pire::Scanner";                    // :271-272

    // Markup(), :381-388.  The image may sit at any alignment in the caller's
    // buffer, so fields are read with memcpy.
    const uint8_t* letters_p = p + pos;
    const uint8_t* final_p = letters_p + (size_t) kMaxChar * 2;
    const uint8_t* final_idx_p = final_p + (size_t) m.final_table_size * 8;
    const uint8_t* rows_p = final_idx_p + (size_t) m.states * 8;

    out->states = m.states;
    out->letters = m.letters;
    out->regexps = m.regexps;
    if (m.initial % row_bytes != 0 || m.initial / row_bytes >= m.states)
        return "scanner image: initial state outside the table";
    out->initial = (uint32_t) (m.initial / row_bytes);                      // StateIndex, :281-284

    out->class_of.resize(kMaxCharUnaligned);
    for (uint32_t c = 0; c < kMaxCharUnaligned; ++c) {
        uint16_t col;
        std::memcpy(&col, letters_p + (size_t) c * 2, 2);
        if (c == 257) {          // Epsilon never reaches a scanner (defs.h:62); slot is 0
            out->class_of[c] = 0;
            continue;
        }
        if (col < header_cells || col >= header_cells + m.letters)
            return "scanner image: letter class out of range";
        out->class_of[c] = (uint16_t) (col - header_cells);
    }

    out->next.resize((size_t) m.states * m.letters);
    out->flags.resize(m.states);
    for (uint32_t s = 0; s < m.states; ++s) {
        const uint8_t* row = rows_p + (size_t) s * row_bytes;
        uint64_t fl;
        std::memcpy(&fl, row + (header_cells == 18 ? 64 : 0), 8);            // Common.Flags
        out->flags[s] = (uint8_t) (fl & 3);
        for (uint32_t c = 0; c < m.letters; ++c) {
            int32_t shift;
            std::memcpy(&shift, row + (size_t) (header_cells + c) * 4, 4);
            // Relocatable::Go: state + SignExtend(shift), :65
            int64_t target = (int64_t) s * (int64_t) row_bytes + shift;
            if (target < 0 || target % (int64_t) row_bytes != 0 || (uint64_t) target / row_bytes >= m.states)
                return "scanner image: transition leaves the table";
            out->next[(size_t) s * m.letters + c] = (uint32_t) ((uint64_t) target / row_bytes);
        }
    }

    out->acc_begin.resize((size_t) m.states + 1);
    for (uint32_t s = 0; s < m.states; ++s) {
        uint64_t at;
        std::memcpy(&at, final_idx_p + (size_t) s * 8, 8);
        out->acc_begin[s] = (uint32_t) out->acc_ids.size();
        for (;; ++at) {
            if (at >= m.final_table_size)
                return "scanner image: accept list not terminated";
            uint64_t id;
            std::memcpy(&id, final_p + at * 8, 8);
            if (id == kEnd)
                break;
            out->acc_ids.push_back((uint32_t) id);
        }
    }
    out->acc_begin[m.states] = (uint32_t) out->acc_ids.size();
    return std::string();
}

} // namespace pire_b200
