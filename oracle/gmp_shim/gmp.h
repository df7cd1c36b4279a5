/*
 * Minimal gmp.h stand-in (TEST INFRASTRUCTURE, not product code).
 *
 * The image ships libgmp.so.10 (GMP 6.3.0) but no development header.  The
 * reference runtime (code_producers/src/c_elements/generic/fr.cpp and
 * common/{main,calcwit}.cpp) includes <gmp.h>; this file declares only the
 * GMP entry points those sources use, with the ABI of GMP 6.x on LP64.  It is
 * used solely by oracle/build_ref.py to compile the reference into
 * oracle/_ref/.
 */
#ifndef CIRCOM_B200_GMP_SHIM_H
#define CIRCOM_B200_GMP_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef long mp_size_t;
typedef unsigned long mp_bitcnt_t;
typedef mp_limb_t *mp_ptr;
typedef const mp_limb_t *mp_srcptr;

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

/* mpn layer */
#define mpn_add_n __gmpn_add_n
#define mpn_sub_n __gmpn_sub_n
#define mpn_add_1 __gmpn_add_1
#define mpn_sub_1 __gmpn_sub_1
#define mpn_add __gmpn_add
#define mpn_cmp __gmpn_cmp
#define mpn_copyi __gmpn_copyi
#define mpn_mul_1 __gmpn_mul_1
#define mpn_addmul_1 __gmpn_addmul_1
#define mpn_zero_p __gmpn_zero_p
#define mpn_and_n __gmpn_and_n
#define mpn_ior_n __gmpn_ior_n
#define mpn_xor_n __gmpn_xor_n
#define mpn_com __gmpn_com
#define mpn_lshift __gmpn_lshift
#define mpn_rshift __gmpn_rshift

mp_limb_t __gmpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_add(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
int __gmpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_copyi(mp_ptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_mul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
int __gmpn_zero_p(mp_srcptr, mp_size_t);
void __gmpn_and_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_ior_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_xor_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_com(mp_ptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_lshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t __gmpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);

/* mpz layer */
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_import __gmpz_import
#define mpz_export __gmpz_export
#define mpz_set_si __gmpz_set_si
#define mpz_set_ui __gmpz_set_ui
#define mpz_add __gmpz_add
#define mpz_sub __gmpz_sub
#define mpz_fits_sint_p __gmpz_fits_sint_p
#define mpz_get_si __gmpz_get_si
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_init_set_si __gmpz_init_set_si
#define mpz_init_set_str __gmpz_init_set_str
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_fdiv_r __gmpz_fdiv_r
#define mpz_fdiv_q __gmpz_fdiv_q
#define mpz_get_str __gmpz_get_str
#define mpz_powm __gmpz_powm
#define mpz_invert __gmpz_invert
/* used by c_elements/goldilocks/fr.hpp:84-106 */
#define mpz_add_ui __gmpz_add_ui
#define mpz_get_ui __gmpz_get_ui
#define mpz_tdiv_q_2exp __gmpz_tdiv_q_2exp

void __gmpz_init(mpz_ptr);
void __gmpz_clear(mpz_ptr);
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *__gmpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
void __gmpz_set_si(mpz_ptr, long);
void __gmpz_set_ui(mpz_ptr, unsigned long);
void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
int __gmpz_fits_sint_p(mpz_srcptr);
long __gmpz_get_si(mpz_srcptr);
void __gmpz_init_set_ui(mpz_ptr, unsigned long);
void __gmpz_init_set_si(mpz_ptr, long);
int __gmpz_init_set_str(mpz_ptr, const char *, int);
size_t __gmpz_sizeinbase(mpz_srcptr, int);
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_fdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
char *__gmpz_get_str(char *, int, mpz_srcptr);
void __gmpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
int __gmpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
unsigned long __gmpz_get_ui(mpz_srcptr);
void __gmpz_tdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);

#ifdef __cplusplus
}
#endif

#endif
