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
	runtime.LockOSThread() // bgn_global_last_error is per OS thread: read it on the thread that failed
	defer runtime.UnlockOSThread()
	if st := C.bgn_ctx_create(&prm, C.int(device), &e.ctx); st != 0 {
		return nil, errors.New("bgn_b200: " + C.GoString(C.bgn_global_last_error()))
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

// ---- chained calls on device-resident batches (include/bgn_b200.h: bgn_buf) ----------------------
// A Batch keeps `count` elements of one group on the GPU in the kernels' own form: a pipeline
// EncryptPoly -> AddPoly -> MultPoly -> AddPoly ... -> DecryptPoly pays Element.SetBytes / Element.Bytes
// (and the curve check of every imported G1 point) once at the edges instead of at every call.
type Batch struct {
	e *Engine
	h *C.bgn_buf
}

func (e *Engine) newBatch(h *C.bgn_buf) *Batch {
	b := &Batch{e, h}
	runtime.SetFinalizer(b, func(b *Batch) { C.bgn_buf_free(b.h) })
	return b
}

// Import is Element.SetBytes over a batch (l2: GT elements, else G1 points).
func (e *Engine) Import(bytes []byte, l2 bool) (*Batch, error) {
	kind := C.int(C.BGN_KIND_G1)
	if l2 {
		kind = C.BGN_KIND_GT
	}
	var h *C.bgn_buf
	st := C.bgn_buf_import(e.ctx, kind, ptr(bytes), C.size_t(len(bytes)/e.ElemBytes), &h)
	return e.newBatch(h), statusErr(e, st)
}

// Bytes is Element.Bytes over the batch.
func (b *Batch) Bytes() ([]byte, error) {
	var kind C.int
	var n C.size_t
	C.bgn_buf_info(b.h, &kind, &n)
	out := make([]byte, int(n)*b.e.ElemBytes)
	return out, statusErr(b.e, C.bgn_buf_export(b.e.ctx, b.h, ptr(out)))
}

func (e *Engine) EncryptPolyBatchH(coeffs []int64, r []byte) (*Batch, error) {
	var h *C.bgn_buf
	st := C.bgn_encrypt_h(e.ctx, (*C.int64_t)(unsafe.Pointer(&coeffs[0])), ptr(r), C.size_t(len(coeffs)), &h)
	return e.newBatch(h), statusErr(e, st)
}

// AddPolyBatchH: level-1 batches (G1 addition, or subtraction) -- 736 M coefficient additions/s on a B200
// against 380 M through the byte format.
func (e *Engine) AddPolyBatchH(a, b *Batch, subtract bool) (*Batch, error) {
	var h *C.bgn_buf
	st := C.bgn_g1_add_h(e.ctx, a.h, b.h, cbool(subtract), &h)
	return e.newBatch(h), statusErr(e, st)
}

func (e *Engine) MultPolyBatchH(c1 *Batch, d1 int, c2 *Batch, d2 int, count int) (*Batch, error) {
	var h *C.bgn_buf
	st := C.bgn_multpoly_h(e.ctx, c1.h, C.size_t(d1), c2.h, C.size_t(d2), C.size_t(count), &h)
	return e.newBatch(h), statusErr(e, st)
}

func (e *Engine) SumL2H(terms *Batch, nterms, ncoeff int) (*Batch, error) {
	var h *C.bgn_buf
	st := C.bgn_l2_sum_reduce_h(e.ctx, terms.h, C.size_t(nterms), C.size_t(ncoeff), &h)
	return e.newBatch(h), statusErr(e, st)
}

func (e *Engine) DecryptBatchH(cts *Batch) ([]int64, []byte, error) {
	var kind C.int
	var n C.size_t
	C.bgn_buf_info(cts.h, &kind, &n)
	vals, status := make([]int64, int(n)), make([]byte, int(n))
	st := C.bgn_decrypt_h(e.ctx, cts.h, (*C.int64_t)(unsafe.Pointer(&vals[0])), ptr(status))
	return vals, status, statusErr(e, st)
}

// ---- several GPUs from one process ---------------------------------------------------------------
// InnerProductMultiGPU shards sum_i u[i]*v[i] by index over one Engine per device, one goroutine each
// (the per-coefficient fan-out of poly.go:129-153 lifted to per-GPU shards), and folds the partial sums
// on the first engine.  tests/cabi/multi_gpu.c is the same program in C with pthreads; on 8 B200s one
// process reaches the throughput of eight (profiles/r02_bench_single_process_n8.json).
func InnerProductMultiGPU(engines []*Engine, u, v []byte, d, count int) ([]byte, error) {
	parts := make([][]byte, len(engines))
	errs := make([]error, len(engines))
	done := make(chan int, len(engines))
	eb := engines[0].ElemBytes
	for g, e := range engines {
		go func(g int, e *Engine) {
			runtime.LockOSThread()
			defer runtime.UnlockOSThread()
			base, rem := count/len(engines), count%len(engines)
			lo := g*base + min(g, rem)
			hi := lo + base
			if g < rem {
				hi++
			}
			parts[g], errs[g] = e.InnerProduct(u[lo*d*eb:hi*d*eb], v[lo*d*eb:hi*d*eb], d, hi-lo)
			done <- g
		}(g, e)
	}
	for range engines {
		<-done
	}
	var all []byte
	for g := range engines {
		if errs[g] != nil {
			return nil, errs[g]
		}
		all = append(all, parts[g]...)
	}
	return engines[0].SumL2(all, len(engines), 2*d)
}
