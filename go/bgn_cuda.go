// bgn_cuda.go -- the cgo binding a maintainer of sachaservan/bgn adds to package bgn to reach
// the B200 engine.  NOT COMPILED IN THIS REPO'S CI: the image has no Go toolchain and no
// libpbc (SURVEY.md 8(c)); the same C-ABI is exercised from Python (tests/) and C++.
//
// Build:  CGO_CFLAGS="-I$REPO/include" CGO_LDFLAGS="-L$REPO/bgn_b200 -lbgn_b200" go build
//
// The file adds the batch entry points north_star names (EncryptPolyBatch, AddPolyBatch,
// MultPolyBatch, DecryptPolyBatch, InnerProduct) beside the existing per-element methods; the
// existing types keep their layout: Ciphertext{C *pbc.Element, L2 bool} (ciphertext.go:12-15),
// PolyCiphertext{Coefficients, Degree, ScaleFactor, L2} (ciphertext.go:26-31).  Elements cross
// the boundary as pbc Element.Bytes() (element_to_bytes): x||y for G1, re||im for GT.
package bgn

/*
#cgo LDFLAGS: -lbgn_b200
#include <stdlib.h>
#include "bgn_b200.h"
*/
import "C"

import (
	"errors"
	"math/big"
	"runtime"
	"unsafe"
)

// Engine is one key resident on one GPU (bgn_ctx).
type Engine struct {
	ctx        *C.bgn_ctx
	ElemBytes  int // 2 * ceil(bits(p)/8)
	ScalarSize int // ceil(bits(n)/8)
}

func statusErr(e *Engine, st C.int) error {
	if st == 0 {
		return nil
	}
	if e != nil && e.ctx != nil {
		return errors.New("bgn_b200: " + C.GoString(C.bgn_last_error(e.ctx)))
	}
	return errors.New("bgn_b200: call failed")
}

// NewEngine uploads the public parameters (bgn.go:28-41) to `device`:
// p, n from pk.PairingParams ("type a1 / p / n / l", parsed as bgn.go:583-593 does for l),
// P and Q as Element.Bytes() (bgn.go:605-607).
func NewEngine(pk *PublicKey, p, n *big.Int, l uint64, device int) (*Engine, error) {
	pb, nb := p.Bytes(), n.Bytes()
	P, Q := pk.P.Bytes(), pk.Q.Bytes()
	prm := C.bgn_params{
		p_be: (*C.uint8_t)(C.CBytes(pb)), p_len: C.size_t(len(pb)),
		n_be: (*C.uint8_t)(C.CBytes(nb)), n_len: C.size_t(len(nb)),
		l:       C.uint64_t(l),
		P_bytes: (*C.uint8_t)(C.CBytes(P)), Q_bytes: (*C.uint8_t)(C.CBytes(Q)),
	}
	defer C.free(unsafe.Pointer(prm.p_be))
	defer C.free(unsafe.Pointer(prm.n_be))
	defer C.free(unsafe.Pointer(prm.P_bytes))
	defer C.free(unsafe.Pointer(prm.Q_bytes))
	e := &Engine{}
	if st := C.bgn_ctx_create(&prm, C.int(device), &e.ctx); st != 0 {
		return nil, errors.New("bgn_b200: bgn_ctx_create failed")
	}
	var limbs, cb, sb C.int
	C.bgn_ctx_info(e.ctx, &limbs, &cb, &sb)
	e.ElemBytes, e.ScalarSize = 2*int(cb), int(sb)
	runtime.SetFinalizer(e, func(e *Engine) { C.bgn_ctx_destroy(e.ctx) })
	return e, nil
}

// SetupDecryption replaces PublicKey.SetupDecryption + PrecomputeTables (bgn.go:195-201, gsbs.go:41-51).
func (e *Engine) SetupDecryption(sk *SecretKey, msgSpace *big.Int) error {
	q := sk.Key.Bytes()
	return statusErr(e, C.bgn_ctx_set_secret(e.ctx, (*C.uint8_t)(unsafe.Pointer(&q[0])), C.size_t(len(q)),
		C.uint64_t(msgSpace.Uint64()), 0))
}

func ptr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&b[0]))
}

// EncryptPolyBatch: coeffs[i] are the polynomial digits of plaintext i (NewPolyPlaintext), all padded to
// `degree` slots; r holds count*degree big-endian scalars of ScalarSize bytes from crypto/rand
// (the randomness stays on the Go side, bgn.go:567-574).  Replaces the loop of poly.go:11-29.
func (e *Engine) EncryptPolyBatch(coeffs []int64, r []byte) ([]byte, error) {
	out := make([]byte, len(coeffs)*e.ElemBytes)
	st := C.bgn_encrypt_batch(e.ctx, (*C.int64_t)(unsafe.Pointer(&coeffs[0])), ptr(r), C.size_t(len(coeffs)), ptr(out))
	return out, statusErr(e, st)
}

// AddPolyBatch: coefficient-wise Add (poly.go:191-204 -> bgn.go:482 / 460) over aligned batches.
func (e *Engine) AddPolyBatch(a, b []byte, l2 bool) ([]byte, error) {
	out := make([]byte, len(a))
	n := C.size_t(len(a) / e.ElemBytes)
	var st C.int
	if l2 {
		st = C.bgn_gt_mul_batch(e.ctx, ptr(a), ptr(b), n, ptr(out))
	} else {
		st = C.bgn_g1_add_batch(e.ctx, ptr(a), ptr(b), n, ptr(out))
	}
	return out, statusErr(e, st)
}

// MultPolyBatch replaces MultPoly (poly.go:123-156) for `count` pairs; out has d1+d2 slots per pair.
func (e *Engine) MultPolyBatch(c1 []byte, d1 int, c2 []byte, d2 int, count int) ([]byte, error) {
	out := make([]byte, count*(d1+d2)*e.ElemBytes)
	st := C.bgn_multpoly_batch(e.ctx, ptr(c1), C.size_t(d1), ptr(c2), C.size_t(d2), C.size_t(count), ptr(out))
	return out, statusErr(e, st)
}

// InnerProduct: AddPoly folded over MultPoly(u[i], v[i]) -- one GPU's share; shares from several
// GPUs are folded with SumL2 again.
func (e *Engine) InnerProduct(u []byte, v []byte, d int, count int) ([]byte, error) {
	prod, err := e.MultPolyBatch(u, d, v, d, count)
	if err != nil {
		return nil, err
	}
	return e.SumL2(prod, count, 2*d)
}

func (e *Engine) SumL2(terms []byte, nterms, ncoeff int) ([]byte, error) {
	out := make([]byte, ncoeff*e.ElemBytes)
	st := C.bgn_l2_sum_reduce(e.ctx, ptr(terms), C.size_t(nterms), C.size_t(ncoeff), ptr(out))
	return out, statusErr(e, st)
}

// DecryptPolyBatch replaces the loop of DecryptPoly (poly.go:32-42) -> decrypt (bgn.go:218-250) ->
// getDL (gsbs.go:54-106).  status[i] == 1 is "cannot find discrete log; out of bounds".
func (e *Engine) DecryptPolyBatch(cts []byte, l2 bool) ([]int64, []byte, error) {
	n := len(cts) / e.ElemBytes
	vals, status := make([]int64, n), make([]byte, n)
	isL2 := C.int(0)
	if l2 {
		isL2 = 1
	}
	st := C.bgn_decrypt_batch(e.ctx, ptr(cts), isL2, C.size_t(n), (*C.int64_t)(unsafe.Pointer(&vals[0])), ptr(status))
	return vals, status, statusErr(e, st)
}

// PolyCiphertextFromBatch rebuilds the reference's container from one polynomial of a batch buffer
// (Element.SetBytes on G1 / GT as NewPolyCiphertextFromBytes does, bgn.go:536-560 -- without the
// per-coefficient Pair(Q,Q) that function performs to obtain a GT-typed element).
func (pk *PublicKey) PolyCiphertextFromBatch(buf []byte, elem int, degree, scaleFactor int, l2 bool) *PolyCiphertext {
	coeffs := make([]*Ciphertext, degree)
	for i := 0; i < degree; i++ {
		var el = pk.G1.NewFieldElement()
		if l2 {
			el = pk.Pairing.NewGT().NewFieldElement()
		}
		el.SetBytes(buf[i*elem : (i+1)*elem])
		coeffs[i] = &Ciphertext{el, l2}
	}
	return &PolyCiphertext{coeffs, degree, scaleFactor, l2}
}

// ---- non-deterministic mode (Deterministic == false) --------------------------------------------
// Blind re-randomises a batch of coefficients the way every homomorphic operation of the reference
// does when !pk.Deterministic: `+ r*Q` on level 1 (bgn.go:264-268, 491-495), `* e(Q,Q)^r` on level 2
// (bgn.go:283-287, 306-310, 469-474).  r: count big-endian scalars of ScalarSize bytes drawn by the
// caller with newCryptoRandom(pk.N) (bgn.go:567-574).  e(Q,Q) is a per-context table, not a pairing
// per call.
func (e *Engine) Blind(cts []byte, r []byte, l2 bool) ([]byte, error) {
	out := make([]byte, len(cts))
	n := C.size_t(len(cts) / e.ElemBytes)
	var st C.int
	if l2 {
		st = C.bgn_gt_blind_batch(e.ctx, ptr(cts), ptr(r), n, ptr(out))
	} else {
		st = C.bgn_g1_blind_batch(e.ctx, ptr(cts), ptr(r), n, ptr(out))
	}
	return out, statusErr(e, st)
}

// MultConstPolyBatch replaces MultConstPoly (poly.go:71-120) for `count` polynomials of `degree` slots:
// digits are the coefficients of pk.NewUnbalancedPlaintext(|constant|) (poly.go:78-80), negate is
// constant < 0 (poly.go:116-118).  The result has degree+len(digits) slots per polynomial.
func (e *Engine) MultConstPolyBatch(cts []byte, degree int, l2 bool, digits []byte, negate bool, count int) ([]byte, error) {
	out := make([]byte, count*(degree+len(digits))*e.ElemBytes)
	st := C.bgn_multconstpoly_batch(e.ctx, ptr(cts), C.size_t(degree), cbool(l2), ptr(digits), C.size_t(len(digits)),
		cbool(negate), C.size_t(count), ptr(out))
	return out, statusErr(e, st)
}

// EvalPolyBatch replaces EvalPoly (poly.go:58-68): one element per polynomial.
func (e *Engine) EvalPolyBatch(cts []byte, degree int, l2 bool, polyBase int, count int) ([]byte, error) {
	out := make([]byte, count*e.ElemBytes)
	st := C.bgn_evalpoly_batch(e.ctx, ptr(cts), C.size_t(degree), cbool(l2), C.uint32_t(polyBase), C.size_t(count), ptr(out))
	return out, statusErr(e, st)
}

// MakePolyL2Batch replaces MakePolyL2 (poly.go:159-163) in deterministic mode: degree+1 slots per polynomial.
func (e *Engine) MakePolyL2Batch(cts []byte, degree int, count int) ([]byte, error) {
	out := make([]byte, count*(degree+1)*e.ElemBytes)
	st := C.bgn_make_poly_l2_batch(e.ctx, ptr(cts), C.size_t(degree), C.size_t(count), ptr(out))
	return out, statusErr(e, st)
}

func cbool(b bool) C.int {
	if b {
		return 1
	}
	return 0
}
