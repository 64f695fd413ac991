// Stand-in for the two Boost.Preprocessor macros the reference's medium-source optimizers use: token pasting and a
// for-each over a sequence (A)(B)(C)... of at most eight elements.
#ifndef CLSIM_REF_SHIM_PREPROCESSOR_HPP
#define CLSIM_REF_SHIM_PREPROCESSOR_HPP
#define BOOST_PP_CAT(a, b) BOOST_PP_CAT_I(a, b)
#define BOOST_PP_CAT_I(a, b) a##b

#define REF_PP_SEQ_HEAD(seq) REF_PP_SEQ_HEAD_1(REF_PP_SEQ_HEAD_0 seq)
#define REF_PP_SEQ_HEAD_0(x) x, REF_PP_NIL
#define REF_PP_SEQ_HEAD_1(...) REF_PP_SEQ_HEAD_2(__VA_ARGS__)
#define REF_PP_SEQ_HEAD_2(x, ...) x
#define REF_PP_SEQ_TAIL(seq) REF_PP_EAT seq
#define REF_PP_EAT(x)

#define REF_PP_SEQ_SIZE(seq) BOOST_PP_CAT(REF_PP_SEQ_SIZE_, REF_PP_SEQ_SIZE_0 seq)
#define REF_PP_SEQ_SIZE_0(_) REF_PP_SEQ_SIZE_1
#define REF_PP_SEQ_SIZE_1(_) REF_PP_SEQ_SIZE_2
#define REF_PP_SEQ_SIZE_2(_) REF_PP_SEQ_SIZE_3
#define REF_PP_SEQ_SIZE_3(_) REF_PP_SEQ_SIZE_4
#define REF_PP_SEQ_SIZE_4(_) REF_PP_SEQ_SIZE_5
#define REF_PP_SEQ_SIZE_5(_) REF_PP_SEQ_SIZE_6
#define REF_PP_SEQ_SIZE_6(_) REF_PP_SEQ_SIZE_7
#define REF_PP_SEQ_SIZE_7(_) REF_PP_SEQ_SIZE_8
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_0 0
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_1 1
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_2 2
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_3 3
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_4 4
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_5 5
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_6 6
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_7 7
#define REF_PP_SEQ_SIZE_REF_PP_SEQ_SIZE_8 8

// SEQ_FOR_EACH(macro, data, (a)(b)(c)) -> macro(r, data, a) macro(r, data, b) macro(r, data, c)
#define BOOST_PP_SEQ_FOR_EACH(macro, data, seq) BOOST_PP_CAT(REF_PP_FE_, REF_PP_SEQ_SIZE(seq))(macro, data, seq)
#define REF_PP_FE_0(macro, data, seq)
#define REF_PP_FE_1(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq))
#define REF_PP_FE_2(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq)) REF_PP_FE_1(macro, data, REF_PP_SEQ_TAIL(seq))
#define REF_PP_FE_3(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq)) REF_PP_FE_2(macro, data, REF_PP_SEQ_TAIL(seq))
#define REF_PP_FE_4(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq)) REF_PP_FE_3(macro, data, REF_PP_SEQ_TAIL(seq))
#define REF_PP_FE_5(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq)) REF_PP_FE_4(macro, data, REF_PP_SEQ_TAIL(seq))
#define REF_PP_FE_6(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq)) REF_PP_FE_5(macro, data, REF_PP_SEQ_TAIL(seq))
#define REF_PP_FE_7(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq)) REF_PP_FE_6(macro, data, REF_PP_SEQ_TAIL(seq))
#define REF_PP_FE_8(macro, data, seq) macro(1, data, REF_PP_SEQ_HEAD(seq)) REF_PP_FE_7(macro, data, REF_PP_SEQ_TAIL(seq))
#endif
