/* oracle/pire_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's scan path over the reference's own
 * serialised scanner image (Scanner::Save, pire/scanners/multi.h:557-573).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may use it.
 * Parity status: PINNED -- checked against the real reference compiled from
 * /root/reference (oracle/_ref/libpire_ref.so) on every ACCEPTS/DENIES vector
 * of tests/pire_ut.cpp that concerns this path and on the committed golden
 * fixtures under tests/golden/ (see tests/test_oracle.py).
 */
#ifndef PIRE_ORACLE_H
#define PIRE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PIRE_ORACLE_BEGIN_MARK = 258,   /* pire/defs.h:63 */
    PIRE_ORACLE_END_MARK   = 259,   /* pire/defs.h:64 */
    PIRE_ORACLE_MAX_CHAR   = 264    /* pire/defs.h:71 */
};

/* A view into a serialised Pire::Scanner (Relocatable).  All pointers alias
 * the caller's blob; the blob must stay alive and 8-byte aligned. */
typedef struct pire_oracle_scanner {
    uint32_t states, letters, regexps, final_table_size;
    uint64_t initial;          /* byte offset of the initial row from transitions */
    uint64_t shortcutting;     /* 0x1000 = NoShortcuts, 0x2000+N = ExitMasks<N> */
    uint32_t header_cells;     /* HEADER_SIZE in 4-byte cells (multi.h:349) */
    uint32_t row_cells;        /* RowSize() (multi.h:347) */
    int      empty;            /* Scanner::Empty() at Save time */
    const uint16_t* letter_of; /* m_letters[264], already + HEADER_SIZE (multi.h:373-375) */
    const uint64_t* final_tab; /* m_final  */
    const uint64_t* final_idx; /* m_finalIndex */
    const uint32_t* trans;     /* m_transitions, Relocatable cells (multi.h:55-69) */
} pire_oracle_scanner;

/* Parse the stream; returns 0 or a negative error (bad magic, truncated ...). */
int pire_oracle_load(const void* blob, size_t size, pire_oracle_scanner* out);

/* State = byte offset of a row from trans (the reference's State minus the
 * m_transitions address). */
uint64_t pire_oracle_initial(const pire_oracle_scanner* sc);                       /* multi.h:161 */
uint64_t pire_oracle_step(const pire_oracle_scanner* sc, uint64_t st, unsigned ch); /* run.h:50-57 */
uint64_t pire_oracle_run(const pire_oracle_scanner* sc, uint64_t st,
                         const uint8_t* begin, const uint8_t* end);                /* run.h:271-275 */
/* Same result as pire_oracle_run, but follows the ExitMasks fast-forward of
 * multi.h:938-1000 (skip 16-byte words holding no exit byte; stop on NoExit). */
uint64_t pire_oracle_run_shortcut(const pire_oracle_scanner* sc, uint64_t st,
                                  const uint8_t* begin, const uint8_t* end);
int      pire_oracle_final(const pire_oracle_scanner* sc, uint64_t st);            /* multi.h:143 */
int      pire_oracle_dead(const pire_oracle_scanner* sc, uint64_t st);             /* multi.h:147 */
uint64_t pire_oracle_state_index(const pire_oracle_scanner* sc, uint64_t st);      /* multi.h:281-284 */
/* AcceptedRegexps (multi.h:149-158): writes up to cap ids, returns the count. */
size_t   pire_oracle_accepted(const pire_oracle_scanner* sc, uint64_t st, uint64_t* ids, size_t cap);

/* Batch driver = RunHelper semantics (run.h:365-392) per string:
 *   Initialize; [Begin()]; Run(str); [End()]; report.
 * offsets == NULL: n fixed-length strings of fixed_len bytes at stride fixed_len.
 * Any of the three outputs may be NULL.  mask_out bit i = regexp id i (< 32). */
void pire_oracle_run_batch(const pire_oracle_scanner* sc, const uint8_t* corpus,
                           const uint64_t* offsets, uint64_t fixed_len, uint64_t n,
                           int with_begin, int with_end, int use_shortcuts,
                           uint8_t* final_out, uint32_t* mask_out, uint32_t* state_out);

/* LongestPrefix / ShortestPrefix (run.h:277-311 with the predicates of run.h:69-100) per
 * string: out[i] = prefix length, or -1 where the reference returns a null pointer. */
void pire_oracle_prefix_batch(const pire_oracle_scanner* sc, const uint8_t* corpus,
                              const uint64_t* offsets, uint64_t fixed_len, uint64_t n,
                              int through_begin, int through_end, int shortest, int64_t* out);

/* LongestSuffix / ShortestSuffix (run.h:316-362), the string walked from its last byte (rbegin) down to its
 * first: out[i] = suffix length, or -1 where the reference returns a null pointer. */
void pire_oracle_suffix_batch(const pire_oracle_scanner* sc, const uint8_t* corpus,
                              const uint64_t* offsets, uint64_t fixed_len, uint64_t n,
                              int through_end, int through_begin, int shortest, int64_t* out);

/* HalfFinalScanner (pire/scanners/half_final.h:136-163) per string, as tests/count_ut.cpp:54-63 drives it:
 *   Initialize (+TakeAction); [Step(BeginMark)]; Step per byte; [Step(EndMark)]; Result(r) for every regexp.
 * counts: n rows of max(1, regexps) u32.  The image is the same Scanner::Save() stream. */
void pire_oracle_count_batch(const pire_oracle_scanner* sc, const uint8_t* corpus,
                             const uint64_t* offsets, uint64_t fixed_len, uint64_t n,
                             int with_begin, int with_end, uint32_t* counts, uint8_t* final_out);

#ifdef __cplusplus
}
#endif
#endif
