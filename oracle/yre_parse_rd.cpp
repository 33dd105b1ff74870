// oracle/yre_parse_rd.cpp -- TEST INFRASTRUCTURE, not product code.
//
// The reference generates its regexp parser with bison from pire/re_parser.y
// (pire/Makefile.am:132-141).  bison is not installed in this image, so the
// oracle build (oracle/build_ref.sh) compiles this hand-written recursive-
// descent driver in its place.  It implements the grammar of
// re_parser.y:82-159 and performs the same Fsm operations per production.
// The two helper functions the grammar's actions call (AppendRange and
// ConvertToFSM, re_parser.y:190-243) are NOT restated here: build_ref.sh
// extracts them verbatim from the read-only reference tree into
// oracle/_ref/gen/re_parser_helpers.inc (git-ignored) at build time.
//
// Gate: the reference's own tests/pire_ut.cpp + tests/easy_ut.cpp must pass
// when linked against this parser (build_ref.sh runs them).
//
// Grammar (re_parser.y):
//   regexp        := alternative                                   :82-90
//   alternative   := conjunction ('|' conjunction)*                :92-95
//   conjunction   := negation (YRE_AND negation)*                  :97-100
//   negation      := [YRE_NOT] concatenation                       :102-105
//   concatenation := iteration*                                    :107-120
//   iteration     := term [YRE_COUNT]                              :122-151
//   term          := LETTERS | DOT | '^' | '$' | '(' alternative ')' :153-159
//
// bison performs default reductions without fetching a look-ahead token; the
// order of Lexer::Lex() calls relative to Lexer::Parenthesized() calls is
// observable by Features (e.g. capture), so the look-ahead below is lazy.

#include <memory>
#include <stdexcept>

#include "fsm.h"
#include "re_lexer.h"
#include "any.h"
#include "stub/stl.h"
#include "re_parser.h"

namespace {

using namespace Pire;
using Pire::Fsm;
using Pire::Encoding;

Fsm& ConvertToFSM(const Encoding& encoding, Any* any);
void AppendRange(const Encoding& encoding, Fsm& a, const Term::CharacterRange& cr);

#include "re_parser_helpers.inc"

struct SyntaxError {};

class Descent {
public:
	explicit Descent(Lexer& lex): L(lex), have(false), tokType(0) {}

	// regexp := alternative <end>
	void Regexp()
	{
		std::unique_ptr<Any> top(Alternative());
		if (Peek() != 0)
			throw SyntaxError();
		ConvertToFSM(L.Encoding(), top.get());
		DoSwap(L.Retval(), *top);
	}

private:
	Lexer& L;
	bool have;
	int tokType;
	std::unique_ptr<Any> tokVal;

	// Token fetch mirrors yylex() of re_parser.y:161-176: a lexer error is
	// remembered in the Lexer and reported as end-of-input.
	void Fetch()
	{
		try {
			Term t = L.Lex();
			tokVal.reset(t.Value().Empty() ? nullptr : new Any(t.Value()));
			tokType = t.Type();
		} catch (Pire::Error& e) {
			L.SetErrMsg(e.what());
			tokVal.reset();
			tokType = 0;
		}
		have = true;
	}
	int Peek() { if (!have) Fetch(); return tokType; }
	Any* Take() { if (!have) Fetch(); have = false; return tokVal.release(); }

	static bool StartsTerm(int t)
	{
		return t == YRE_LETTERS || t == YRE_DOT || t == '^' || t == '$' || t == '(';
	}

	Any* Alternative()
	{
		std::unique_ptr<Any> lhs(Conjunction());
		while (Peek() == '|') {
			delete Take();
			std::unique_ptr<Any> rhs(Conjunction());
			ConvertToFSM(L.Encoding(), lhs.get()) |= ConvertToFSM(L.Encoding(), rhs.get());
		}
		return lhs.release();
	}

	Any* Conjunction()
	{
		std::unique_ptr<Any> lhs(Negation());
		while (Peek() == YRE_AND) {
			delete Take();
			std::unique_ptr<Any> rhs(Negation());
			ConvertToFSM(L.Encoding(), lhs.get()) &= ConvertToFSM(L.Encoding(), rhs.get());
		}
		return lhs.release();
	}

	Any* Negation()
	{
		if (Peek() == YRE_NOT) {
			delete Take();
			std::unique_ptr<Any> body(Concatenation());
			ConvertToFSM(L.Encoding(), body.get()).Complement();
			return body.release();
		}
		return Concatenation();
	}

	Any* Concatenation()
	{
		std::unique_ptr<Any> acc(new Any(Fsm()));
		while (StartsTerm(Peek())) {
			std::unique_ptr<Any> piece(Iteration());
			Fsm& a = ConvertToFSM(L.Encoding(), acc.get());
			if (piece->IsA<Term::CharacterRange>() && !piece->As<Term::CharacterRange>().second)
				AppendRange(L.Encoding(), a, piece->As<Term::CharacterRange>());
			else if (piece->IsA<Term::DotTag>())
				L.Encoding().AppendDot(a);
			else
				a += ConvertToFSM(L.Encoding(), piece.get());
		}
		return acc.release();
	}

	Any* Iteration()
	{
		std::unique_ptr<Any> base(TermRule());
		if (Peek() != YRE_COUNT)
			return base.release();

		std::unique_ptr<Any> cnt(Take());
		const Term::RepetitionCount& rep = cnt->As<Term::RepetitionCount>();
		Fsm& orig = ConvertToFSM(L.Encoding(), base.get());
		std::unique_ptr<Any> out(new Any(orig));
		Fsm& cur = out->As<Fsm>();

		const int lo = rep.first, hi = rep.second;
		if (lo == 0 && hi == 1) {
			Fsm nothing;
			cur |= nothing;
		} else if (lo == 0 && hi == Inf) {
			cur.Iterate();
		} else if (lo == 1 && hi == Inf) {
			cur += *cur;
		} else {
			cur *= lo;
			if (hi == Inf)
				cur += *orig;
			else if (hi != lo)
				cur += (orig | Fsm()) * (hi - lo);
		}
		L.Parenthesized(out->As<Fsm>());
		return out.release();
	}

	Any* TermRule()
	{
		int t = Peek();
		if (t == '(') {
			delete Take();
			std::unique_ptr<Any> inner(Alternative());
			if (Peek() != ')')
				throw SyntaxError();
			// bison reduces '(' alternative ')' as soon as ')' is shifted,
			// before asking the lexer for anything else.
			std::unique_ptr<Any> close(Take());
			L.Parenthesized(inner->As<Fsm>());
			return inner.release();
		}
		if (!StartsTerm(t))
			throw SyntaxError();
		return Take();
	}
};

} // namespace

namespace Pire {
namespace Impl {

	// Same contract as re_parser.y:264-275: 0 on success, non-zero on a syntax
	// error; a message left in the Lexer is rethrown as Pire::Error.
	int yre_parse(Pire::Lexer& rlex)
	{
		int rc = 0;
		try {
			Descent(rlex).Regexp();
		} catch (SyntaxError&) {
			rc = 1;
		}
		if (!rlex.ErrMsg().empty())
			throw Error(rlex.ErrMsg());
		return rc;
	}

}
}
