// tests/cpp/mirror_check.cpp -- the reference's own run.h templates instantiated on
// Pire::Gpu::Scanner (host concept over the flattened tables), compared with the
// same templates on the reference scanner.  Built and run by tests/test_cpp_mirror.py
// where the reference headers exist.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <pire.h>

#include "pire_gpu.hpp"

template <class Sc>
typename Sc::State RunRegexp(const Sc& sc, const std::string& s)          // tests/common.h:158-169
{
    typename Sc::State st;
    sc.Initialize(st);
    Pire::Step(sc, st, Pire::BeginMark);
    Pire::Run(sc, st, s.c_str(), s.c_str() + s.size());
    Pire::Step(sc, st, Pire::EndMark);
    return st;
}

int main()
{
    const char* patterns[] = {"hello\\s+w.+d$", "abc|def", "^x{3,6}$", "a.{3,10}$", "[ab]{3}"};
    const char* texts[] = {"hello world", "xx hello\tworld", "Hello world", "abc", "deb", "xxx", "xxxxxxx",
                           "bbbbbbbbxeeee", "......aab.....", "", "a", "xaeeeeeeeeeeeeeeee"};
    int bad = 0, checked = 0;
    Pire::Scanner glued;
    for (const char* p : patterns) {
        Pire::Fsm fsm = Pire::Lexer(p).Parse();
        fsm.Surround();
        Pire::Scanner ref = fsm.Compile<Pire::Scanner>();
        glued = glued.Empty() ? ref : Pire::Scanner::Glue(glued, ref);
        Pire::Gpu::Scanner mine(ref, /*device*/ -1);
        if (mine.Size() != ref.Size() || mine.RegexpsCount() != ref.RegexpsCount() || mine.LettersCount() != ref.LettersCount())
            ++bad;
        for (const char* t : texts) {
            Pire::Scanner::State a = RunRegexp(ref, t);
            Pire::Gpu::Scanner::State b = RunRegexp(mine, t);
            ++checked;
            if (ref.StateIndex(a) != mine.StateIndex(b) || ref.Final(a) != mine.Final(b) || ref.Dead(a) != mine.Dead(b))
                ++bad;
            // the fluent form, run.h:365-392
            bool m1 = Pire::Runner(ref).Begin().Run(t, std::strlen(t)).End();
            bool m2 = Pire::Runner(mine).Begin().Run(t, std::strlen(t)).End();
            if (m1 != m2)
                ++bad;
            // prefix scans of run.h:277-311 compile and agree too
            const char* e = t + std::strlen(t);
            if (Pire::LongestPrefix(ref, t, e) != Pire::LongestPrefix(mine, t, e))
                ++bad;
            if (Pire::ShortestPrefix(ref, t, e) != Pire::ShortestPrefix(mine, t, e))
                ++bad;
        }
    }
    // accept lists of a glued scanner, multi.h:149-158
    Pire::Gpu::Scanner mine(glued, -1);
    for (const char* t : texts) {
        auto a = glued.AcceptedRegexps(RunRegexp(glued, t));
        auto b = mine.AcceptedRegexps(RunRegexp(mine, t));
        ++checked;
        if (a.second - a.first != b.second - b.first)
            ++bad;
        else
            for (ptrdiff_t i = 0; i < a.second - a.first; ++i)
                bad += a.first[i] != b.first[i];
    }
    // a host-only handle must refuse to scan (no CPU fallback)
    try {
        uint64_t offs[2] = {0, 3};
        std::vector<bool> out;
        Pire::Gpu::MatchesHost(mine, (const uint8_t*) "abc", offs, 1, out);
        ++bad;
    } catch (Pire::Gpu::Error& e) {
        if (e.Code != PIRE_GPU_ENODEVICE)
            ++bad;
    }
    // a HalfFinalScanner (pire/scanners/half_final.h) ingests through the same Save() stream; the host concept
    // walks the same states, and the counting entry point refuses a host-only handle as well
    {
        Pire::Fsm fsm = Pire::Lexer("ab+c|b").Parse();
        Pire::HalfFinalScanner hf(fsm);
        Pire::Gpu::Scanner hmine(hf, -1);
        ++checked;
        if (hmine.Size() != hf.Size() || hmine.RegexpsCount() != hf.RegexpsCount())
            ++bad;
        for (const char* t : {"abbbc", "b", "xabcb", ""}) {
            Pire::HalfFinalScanner::State a = RunRegexp(hf, t);
            Pire::Gpu::Scanner::State b = RunRegexp(hmine, t);
            ++checked;
            if (hf.StateIndex(a) != hmine.StateIndex(b) || hf.Final(a) != hmine.Final(b))
                ++bad;
        }
        try {
            uint32_t counts[1];
            Pire::Gpu::Batch batch{(const uint8_t*) "abc", nullptr, 3, 1};
            Pire::Gpu::HalfFinalCount(hmine, batch, counts);
            ++bad;
        } catch (Pire::Gpu::Error& e) {
            if (e.Code != PIRE_GPU_ENODEVICE)
                ++bad;
        }
    }
    std::printf("mirror_check: %d comparisons, %d mismatches\n", checked, bad);
    return bad ? 1 : 0;
}
