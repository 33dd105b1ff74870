// pire_image.hpp -- host-side ingest of a compiled Pire scanner.
//
// The drop-in boundary for compiled automata is the byte stream written by
// Pire::Scanner::Save() (reference pire/scanners/multi.h:557-573; layout
// multi.h:297-305,:315-323,:381-388; stream header pire/scanners/common.h:44-78).
// The reference's regex front end (Lexer -> Fsm -> Compile/Glue) stays on the
// host unchanged; this file turns its output into a neutral DFA description
// (state indices instead of row addresses) from which the device tables are
// built (dfa_tables.hpp).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace pire_b200 {

// Alphabet constants of the reference (pire/defs.h:59-73).
constexpr uint32_t kBeginMark = 258;
constexpr uint32_t kEndMark = 259;
constexpr uint32_t kMaxCharUnaligned = 260;
constexpr uint32_t kMaxChar = 264;

// A compiled scanner in index space.
struct Dfa {
    uint32_t states = 0;       // Scanner::Size()
    uint32_t letters = 0;      // Scanner::LettersCount()
    uint32_t regexps = 0;      // Scanner::RegexpsCount()
    uint32_t initial = 0;      // StateIndex(Initialize())
    bool empty = false;        // Scanner::Empty(): never matches anything (multi.h:121,:135)
    bool exit_masks = false;   // saved with ExitMasks<2> (else NoShortcuts)

    // letter class (0..letters-1) of every input symbol 0..259: bytes, Epsilon
    // (unused), BeginMark, EndMark.  Translate() minus HEADER_SIZE (multi.h:163-166,:375).
    std::vector<uint16_t> class_of;          // [kMaxCharUnaligned]
    // next[s * letters + c] = state index after reading a symbol of class c
    // (NextTranslated, multi.h:169-186, with Relocatable::Go :65 resolved).
    std::vector<uint32_t> next;
    // bit0 Final, bit1 Dead (multi.h:90-94,:143-147).
    std::vector<uint8_t> flags;
    // AcceptedRegexps(s) (multi.h:149-158) as CSR: ids [acc_begin[s], acc_begin[s+1]).
    std::vector<uint32_t> acc_begin;
    std::vector<uint32_t> acc_ids;

    uint32_t Next(uint32_t s, uint32_t symbol) const { return next[(size_t) s * letters + class_of[symbol]]; }
    bool Final(uint32_t s) const { return (flags[s] & 1) != 0; }
    bool Dead(uint32_t s) const { return (flags[s] & 2) != 0; }
};

// Parses the stream.  Returns an empty string on success, else a message.
// Rejects what Scanner::Load/Mmap reject (common.h:65-78, multi.h:252-261):
// bad magic/version/pointer size, wrong scanner type, truncated image; and, in
// addition, transitions that leave the table (a corrupted image must not turn
// into out-of-bounds device reads).
std::string ParsePireImage(const void* data, size_t size, Dfa* out);

} // namespace pire_b200
