// pbc_vectors -- closes the parity loop of SURVEY.md 8(f4): recomputes every "out" of a
// tests/golden/kb*.json fixture with the REAL libpbc (through github.com/Nik-U/pbc, the module the
// reference pins in go.mod:5) and writes the result next to it as tests/golden/pbc/kb*.json.
// tests/test_pbc_vectors.py then compares the two files section by section; while no PBC-made file
// exists the test is skipped and parity stays "unpinned" (DESIGN.md 0).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain, no libpbc, no gmp.h).  Build where the
// reference builds:
//
//	cd tools/pbc_vectors && go mod init pbcvectors && go get github.com/Nik-U/pbc@v0.0.0-20181205041846-3e516ca0c5d6
//	go run . ../../tests/golden/kb512.json ../../tests/golden/pbc/kb512.json
//
// Conventions shared with the fixtures: elements are hex of Element.Bytes() (G1 = x||y, GT = re||im);
// the point at infinity is all-zero bytes (PBC itself has no canonical encoding for O: SURVEY.md 8(a)),
// so O is mapped explicitly on the way in (Set0) and on the way out (Is0).
package main

import (
	"bytes"
	"encoding/gob"
	"encoding/hex"
	"encoding/json"
	"fmt"
	"math/big"
	"os"
	"strings"

	"github.com/Nik-U/pbc"
)

type obj = map[string]interface{}

// the reference's wire structs (ciphertext.go:17-20, 33-38), restated so that this tool can emit the
// gob envelopes bgn_b200/gobwire.py reads and writes
type ciphertextWrapper struct {
	CBytes []byte
	L2     bool
}

type polyCiphertextWrapper struct {
	CoeffBytes  [][]byte
	Degree      int
	ScaleFactor int
	L2          bool
}

func gobHex(v interface{}) string {
	var buf bytes.Buffer
	die(gob.NewEncoder(&buf).Encode(v))
	return hex.EncodeToString(buf.Bytes())
}

var pairing *pbc.Pairing
var elemBytes int

func die(err error) {
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
}

func allZero(b []byte) bool {
	for _, x := range b {
		if x != 0 {
			return false
		}
	}
	return true
}

func unhex(s string) []byte {
	b, err := hex.DecodeString(s)
	die(err)
	return b
}

func g1(s string) *pbc.Element {
	b := unhex(s)
	e := pairing.NewG1()
	if allZero(b) {
		return e.Set0()
	}
	return e.SetBytes(b)
}

func gt(s string) *pbc.Element { return pairing.NewGT().SetBytes(unhex(s)) }

func g1hex(e *pbc.Element) string {
	if e.Is0() {
		return hex.EncodeToString(make([]byte, elemBytes))
	}
	return hex.EncodeToString(e.Bytes())
}

func gthex(e *pbc.Element) string { return hex.EncodeToString(e.Bytes()) }

func bigOf(v interface{}) *big.Int {
	switch t := v.(type) {
	case string:
		z, ok := new(big.Int).SetString(strings.TrimPrefix(t, "0x"), 16)
		if !ok {
			die(fmt.Errorf("bad hex integer %q", t))
		}
		return z
	case float64:
		return big.NewInt(int64(t))
	}
	die(fmt.Errorf("bad integer %v", v))
	return nil
}

func strs(v interface{}) []string {
	a := v.([]interface{})
	out := make([]string, len(a))
	for i, x := range a {
		out[i] = x.(string)
	}
	return out
}

// x*P + r*Q with the sign convention of EncryptPoly (poly.go:17-21): x < 0 -> -(|x|P + rQ)
func encrypt(P, Q *pbc.Element, x, r *big.Int) *pbc.Element {
	ax := new(big.Int).Abs(x)
	c := pairing.NewG1().PowBig(P, ax) // bgn.go:344
	h := pairing.NewG1().PowBig(Q, r)  // bgn.go:346
	c = pairing.NewG1().Mul(c, h)      // bgn.go:350
	if x.Sign() < 0 {
		c = pairing.NewG1().Neg(c)
	}
	return c
}

func main() {
	if len(os.Args) != 3 {
		die(fmt.Errorf("usage: pbc_vectors <fixture.json> <out.json>"))
	}
	raw, err := os.ReadFile(os.Args[1])
	die(err)
	var fx obj
	die(json.Unmarshal(raw, &fx))
	pairing, err = pbc.NewPairingFromString(fx["pbc_params"].(string)) // bgn.go:640
	die(err)
	elemBytes = 2 * int(fx["coord_bytes"].(float64))
	P, Q := g1(fx["P"].(string)), g1(fx["Q"].(string))
	q1 := bigOf(fx["q1"])
	out := obj{"note": "produced by libpbc via github.com/Nik-U/pbc (tools/pbc_vectors/main.go)", "key_bits": fx["key_bits"]}

	sec := func(name string) obj { return fx[name].(obj) }
	binop := func(name string, f func(a, b string) string) {
		v := sec(name)
		a, b := strs(v["a"]), strs(v["b"])
		res := make([]string, len(a))
		for i := range a {
			res[i] = f(a[i], b[i])
		}
		out[name] = obj{"out": res}
	}
	unop := func(name string, f func(a string) string) {
		v := sec(name)
		a := strs(v["a"])
		res := make([]string, len(a))
		for i := range a {
			res[i] = f(a[i])
		}
		out[name] = obj{"out": res}
	}

	{ // Encrypt
		v := sec("encrypt")
		xs, rs := v["x"].([]interface{}), v["r"].([]interface{})
		res := make([]string, len(xs))
		for i := range xs {
			res[i] = g1hex(encrypt(P, Q, bigOf(xs[i]), bigOf(rs[i])))
		}
		out["encrypt"] = obj{"out": res}
	}
	binop("g1_add", func(a, b string) string { return g1hex(pairing.NewG1().Mul(g1(a), g1(b))) })  // bgn.go:482
	binop("g1_sub", func(a, b string) string { return g1hex(pairing.NewG1().Div(g1(a), g1(b))) })  // bgn.go:419
	unop("g1_neg", func(a string) string { return g1hex(pairing.NewG1().Neg(g1(a))) })
	{ // MultConst on level 1 (bgn.go:258)
		v := sec("g1_mulconst")
		a, k := strs(v["a"]), v["k"].([]interface{})
		res := make([]string, len(a))
		for i := range a {
			res[i] = g1hex(pairing.NewG1().PowBig(g1(a[i]), bigOf(k[i])))
		}
		out["g1_mulconst"] = obj{"out": res}
	}
	binop("pair", func(a, b string) string { return gthex(pairing.NewGT().Pair(g1(a), g1(b))) }) // bgn.go:300
	unop("make_l2", func(a string) string { return gthex(pairing.NewGT().Pair(g1(a), P)) })      // bgn.go:318
	binop("gt_mul", func(a, b string) string { return gthex(pairing.NewGT().Mul(gt(a), gt(b))) }) // bgn.go:460
	binop("gt_div", func(a, b string) string { return gthex(pairing.NewGT().Div(gt(a), gt(b))) }) // bgn.go:397
	unop("gt_inv", func(a string) string { return gthex(pairing.NewGT().Invert(gt(a))) })
	{ // MultConst on level 2 (bgn.go:277)
		v := sec("gt_pow")
		a, k := strs(v["a"]), v["k"].([]interface{})
		res := make([]string, len(a))
		for i := range a {
			res[i] = gthex(pairing.NewGT().PowBig(gt(a[i]), bigOf(k[i])))
		}
		out["gt_pow"] = obj{"out": res}
	}
	{ // MultPoly (poly.go:123-156)
		v := sec("multpoly")
		c1, c2 := strs(v["c1"]), strs(v["c2"])
		res := make([]*pbc.Element, len(c1)+len(c2))
		for i := range res {
			res[i] = pairing.NewGT().Set1()
		}
		for i := range c1 {
			for k := range c2 {
				e := pairing.NewGT().Pair(g1(c1[i]), g1(c2[k]))
				res[i+k] = pairing.NewGT().Mul(res[i+k], e)
			}
		}
		hx := make([]string, len(res))
		for i := range res {
			hx[i] = gthex(res[i])
		}
		out["multpoly"] = obj{"out": hx}
	}
	{ // C^q1 (bgn.go:223)
		v := sec("decrypt_l2")
		in := strs(v["in"])
		res := make([]string, len(in))
		for i := range in {
			res[i] = gthex(pairing.NewGT().PowBig(gt(in[i]), q1))
		}
		out["decrypt_l2"] = obj{"csk": res}
	}
	{ // re-randomisation of the non-deterministic mode (bgn.go:264-268, 283-287)
		v := sec("g1_blind")
		a, r := strs(v["a"]), v["r"].([]interface{})
		res := make([]string, len(a))
		for i := range a {
			h := pairing.NewG1().PowBig(Q, bigOf(r[i]))
			res[i] = g1hex(pairing.NewG1().Mul(g1(a[i]), h))
		}
		out["g1_blind"] = obj{"out": res}
		v = sec("gt_blind")
		a, r = strs(v["a"]), v["r"].([]interface{})
		res = make([]string, len(a))
		qq := pairing.NewGT().Pair(Q, Q)
		for i := range a {
			h := pairing.NewGT().PowBig(qq, bigOf(r[i]))
			res[i] = gthex(pairing.NewGT().Mul(gt(a[i]), h))
		}
		out["gt_blind"] = obj{"out": res}
	}
	{ // gob envelopes made by Go itself (ciphertext.go:76-116): pins bgn_b200/gobwire.py
		v := sec("multpoly")
		c1 := strs(v["c1"])
		coeffs := make([][]byte, len(c1))
		for i := range c1 {
			coeffs[i] = unhex(c1[i])
		}
		out["gob"] = obj{
			"ciphertext_l1":   gobHex(ciphertextWrapper{CBytes: coeffs[0], L2: false}),
			"ciphertext_l2":   gobHex(ciphertextWrapper{CBytes: unhex(strs(sec("pair")["out"])[0]), L2: true}),
			"poly_ciphertext": gobHex(polyCiphertextWrapper{CoeffBytes: coeffs, Degree: len(coeffs), ScaleFactor: 2, L2: false}),
		}
	}
	enc, err := json.MarshalIndent(out, "", " ")
	die(err)
	die(os.WriteFile(os.Args[2], enc, 0o644))
}
