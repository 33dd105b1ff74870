"""Known-answer vectors of the reference's own unit tests for the scan path.

Transcribed from /root/reference/tests/pire_ut.cpp (ACCEPTS / DENIES macros,
tests/common.h:195-221).  Each ACCEPTS(str) there asserts that
``Initialize; Step(BeginMark); Run(str); Step(EndMark)`` (tests/common.h:158-169)
ends in a state with a non-empty AcceptedRegexps list; DENIES asserts the
opposite.  Patterns are compiled like tests/common.h:40-69 (ParseRegexp): the
option letters are 'i' CaseInsensitive, 'u' UTF-8, 'n' no Surround(), 'a' AndNot.

Only tests expressible as (pattern, options) are listed; the Fsm-algebra cases
(Misc's ``& ~``, Reverse) are compile-time features outside the scan path.
"""

# (test name @ line in pire_ut.cpp, pattern, options, accepts, denies)
VECTORS = [
    ("String@38", b"abc", "", [b"def abc ghi", b"abc"], [b"def abd ghi"]),
    ("Boundaries@47a", b"^abc", "", [b"abc ghi"], [b"def abc"]),
    ("Boundaries@47b", b"abc$", "", [b"def abc"], [b"abc ghi"]),
    ("Primitives@60a", b"abc|def", "", [b"def", b"abc"], [b"deb"]),
    ("Primitives@60b", b"ad*e", "", [b"xaez", b"xadez", b"xaddez", b"xadddddddddddddddddddddddez"], [b"xafez"]),
    ("Primitives@60c", b"ad+e", "", [b"xadez", b"xaddez", b"xadddddddddddddddddddddddez"], [b"xaez", b"xafez"]),
    ("Primitives@60d", b"ad?e", "", [b"xaez", b"xadez"], [b"xaddez", b"xafez"]),
    ("Primitives@60e", b"a.{1}e", "", [b"axe"], [b"ae", b"axye"]),
    ("MassAlternatives@109a", b"((abc|def)|ghi)|klm", "", [b"abc", b"def", b"ghi", b"klm"], [b"aei", b"klc"]),
    ("MassAlternatives@109b", b"(abc|def)|(ghi|klm)", "", [b"abc", b"def", b"ghi", b"klm"], [b"aei", b"klc"]),
    ("MassAlternatives@109c", b"abc|(def|(ghi|klm))", "", [b"abc", b"def", b"ghi", b"klm"], [b"aei", b"klc"]),
    ("MassAlternatives@109d", b"abc|(def|ghi)|klm", "", [b"abc", b"def", b"ghi", b"klm"], [b"aei", b"klc"]),
    ("Composition@120a", rb"^/([^\\/]|\\.)*/[a-z]*$", "",
     [b"/regexp/i", b"/regexp2/", b"/dir\\/file/", b"/dir\\\\/"],
     [b"regexp", b"/dir/file/", b"/dir\\\\/file/"]),
    ("Composition@120b", b"Head(Inner)*Tail", "",
     [b"HeadInnerTail", b"HeadInnerInnerTail", b"HeadTail"], [b"HeadInneInnerTail"]),
    ("Repetition@142a", b"^x{3,6}$", "", [b"xxx", b"xxxx", b"xxxxx", b"xxxxxx"], [b"xx", b"xxxxxxx"]),
    ("Repetition@142b", b"^x{3,}$", "", [b"xxx", b"xxxx", b"x" * 11, b"x" * 47], [b"xx"]),
    ("Repetition@142c", b"^x{3}$", "", [b"xxx"], [b"x", b"xx", b"xxxx", b"xxxxx", b"x" * 47]),
    ("Repetition@142d", b"x.{3,10}$", "",
     [b"b" * (2 * n) + b"x" + b"e" * n for n in range(20) if 3 <= n <= 10],
     [b"b" * (2 * n) + b"x" + b"e" * n for n in range(20) if not 3 <= n <= 10]),
    ("UTF8@181a", b"^.$", "u",
     [b"\x41", b"\xC1\x81", b"\xE1\x81\x82", b"\xF1\x81\x82\x83"],
     [b"\x81", b"\xC1", b"\xC1\x41", b"\xC1\xC2", b"\xC1\x81\x82", b"\xE1", b"\xE1\x42", b"\xE1\x42\x43",
      b"\xE1\xC2\xC3", b"\xE1\x82", b"\xE1\x82\x83\x84"]),
    ("UTF8@181b", b"x\xD0\xA4y", "u", [b"x\xD0\xA4y"], []),
    ("AndNot@211a", b"<([0-9]+&~123&~456)>", "a", [b"<111>", b"<124>"], [b"<123>", b"<456>", b"<abc>"]),
    ("AndNot@211b", rb"[0-9]+\&1+", "a", [b"123&111"], [b"111"]),
    ("Misc@238a", rb"^[^\s=/>]*$", "n", [b"a"], []),
    ("Misc@238b", rb"\t", "", [b"\t"], []),
    ("Ranges@251", rb"a\W", "", [b"a,"], [b"ab"]),
    ("TestShortcuts@628a", b"aaa", "",
     [b"." * 38 + b"aaa" + b"." * 13], [b"." * 38 + b"aab" + b"." * 13, b"." * 54]),
    ("TestShortcuts@628b", b"[ab]{3}", "",
     [b"." * 38 + b"aaa" + b"." * 13, b"." * 38 + b"aab" + b"." * 13, b"." * 38 + b"bbb" + b"." * 13], [b"." * 54]),
    ("TestShortcuts@628c", b"\xD0\xB0", "u",
     [b"." * 38 + b"\xD0\xB0" + b"." * 15, b"." * 35 + b"\xD0\xB0" + b"." * 18, b"." * 32 + b"\xD0\xB0" + b"." * 21], []),
    ("Aligned@729a", b"xy", "", [b"xy"], [b"yz"]),
    ("Aligned@729b", b"abcde", "",
     [b"ZZZZZabcdeZZZZZZ", b"ZabcdeZZZ", b"ZZZZZZZZZZZZZabcde"],
     [b"ZZZZZabcdfZZZZZZ", b"ZxbcdeZZZ", b"ZZZZZZZZZZZZZabcdf"]),
    # README:58-70 -- "Hello world" against the case-insensitive UTF-8 headline regexp
    ("README@58", rb"hello\s+w.+d$", "iu", [b"Hello world", b"hello  w..d"], [b"hello wd", b"hello world!"]),
    # SURVEY.md Appendix A answers for the headline regexp (Latin-1, case-sensitive)
    ("AppendixA", rb"hello\s+w.+d$", "",
     [b"hello world", b"hello  w..d", b"xx hello\tworld"], [b"Hello world", b"hello wd", b"hello world!", b""]),
    # pire_ut.cpp:265-271 (the scanner of the reversed automaton) and :553-579 (the vectors every scanner type
    # must still answer after Save/Load)
    ("Reverse@265", b"abcdef", "r", [b"fedcba"], [b"abcdef"]),
    ("Serialization@553", b"^regexp$", "", [b"regexp"], [b"regxp", b"regexp t"]),
]

# Aligned@729: the same strings are also run at unaligned addresses there; the
# parity tests place every vector at several byte offsets inside the corpus.

# TestGlue@648-693: glue "aaa" + "bbb", then "ccc" in front; exact accept-id lists.
GLUE_CASES = [
    # (patterns glued left to right as (pattern, opts), [(string, expected ids)])
    ([(b"aaa", ""), (b"bbb", "")],
     [(b"aaa", [0]), (b"bbb", [1]), (b"aaabbb", [0, 1]), (b"ccc", [])]),
    ([(b"ccc", ""), (b"aaa", ""), (b"bbb", "")],      # Glue(sc3, glued): ids shift by one
     [(b"ccc", [0]), (b"aaa", [1]), (b"bbb", [2]), (b"aaabbbccc", [0, 1, 2])]),
    ([(b"a", "n"), (b"c", "n")],
     [(b"ac", [])]),
]

# ScanBoundaries@343 (pire_ut.cpp:351-464): (pattern, text, ShortestPrefix length, LongestPrefix length),
# -1 = null.  Patterns are compiled WITHOUT Surround() (pire_ut.cpp:313-325: Lexer(p).Parse().Compile<Scanner>()),
# no Begin/End marks.
_D = b"123456789-" * 8
PREFIX_CASES = [
    (b"a*", b"", 0, 0),
    (b"a", b"", -1, -1),
    (b"fixed", b"fixed prefix", 5, 5),
    (b"fixed", b"a fixed nonexistent prefix", -1, -1),
    (b"a*", b"aaabbb", 0, 3),
    (b"a*", b"bbbbbb", 0, 0),
    (b"a*", b"aaaaaa", 0, 6),
    (b"aa*", b"aaabbb", 1, 3),
    (b"a*a", b"aaaaaa", 1, 6),
    (b".*a", b"bbbba", 5, 5),
    (b".*", _D, 0, 80),
    (b".*a", _D + b"a", 81, 81),
    (b".*a", _D + b"a" + _D + b"a", 81, 162),
    (b".*b", _D, -1, -1),
    (b".*a.*", _D + b"a" + _D + b"b", 81, 162),
    (b".*a.*b", _D + b"a" + _D + b"b", 162, 162),
    (b"1.*a.*", _D + b"a" + _D + b"b", 81, 162),
    (b"a+", b"bbbbbb", -1, -1),
    # ScanTermination@475: "aaa" over "aaab\0": the scan must stop in the dead state; longest prefix = 3
    (b"aaa", b"aaab\0", 3, 3),
]


# tests/count_ut.cpp:540-551 (TestHalfFinalCount) and :562-576 (TestHalfFinalSerialization): pattern (UTF-8 lexer, no
# Surround), text, and Result(0) of the five HalfFinalFsm counters in the order of MakeHalfFinalCount (:503-520):
# MakeGreedyCounter(true), MakeGreedyCounter(false), MakeNonGreedyCounter(true,true), (true,false), (false).
# The glued scanner of all five must report the same numbers as Result(0..4) (:533-537).
COUNT_CASES = [
    (b"ab+", b"abbabbbabbbbbb", [3, 3, 3, 11, 3]),
    (b"(ab)+", b"ababbababbab", [3, 3, 5, 5, 5]),
    (b"(abab)+", b"ababababab", [1, 1, 4, 4, 2]),
    (b"ab+c|b", b"abbbbbbbbbb", [1, 10, 10, 10, 10]),
    (b"ab+c|b", b"abbbbbbbbbbb", [1, 10, 11, 11, 11]),
    (b"ab+c|b", b"abbbbbbbbbbc", [1, 1, 10, 11, 10]),
    (b"ab+c|b", b"abbbbbbbbbbbc", [1, 1, 11, 12, 11]),
    (b"a\\w+c|b", b"abbbdbbbdbbc", [1, 1, 8, 9, 8]),
    (b"a\\w+c|b", b"abbbdbbbdbb", [1, 8, 8, 8, 8]),
    (b"a[a-z]+c|b", b"abeeeebeeeeeeeeeceeaeebeeeaeecceebeeaeebeeb", [2, 4, 7, 9, 7]),
    (b"(\\w\\w)+", b"ab abbb ababa a", [3, 3, 8, 8, 5]),
]
