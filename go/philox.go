// philox.go - Philox4x32-10 counter-based streams of the "--rng=philox" specification (DESIGN.md section 4.1), for a Go
// host that wants the GPU's exact fragment set.  Restates oracle/philox.h (the plain-C text of the spec); the CUDA side is
// rlsim_b200/csrc/philox.cuh.  Known answers (Random123): counter 0 / key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8;
// all ones -> 408f276d 41c83b0e a20bc7c6 6d5451fd; counter 243f6a88 85a308d3 13198a2e 03707344, key a4093822 299f31d0 ->
// d16cfe09 94fdcceb 5001e420 24126ea1 (tests/test_gpu_parity.py::test_philox_blocks checks the same vectors on the GPU).
//
// Not compiled in this repository (no Go toolchain in the build image); drop into the reference's src/ next to
// gpu_pool.go.
package main

import "math/bits"

// PxBlock is one Philox output block: four 32-bit words = two 64-bit draws.
type PxBlock [4]uint32

// PxKey is a stream key (two words).
type PxKey struct{ K0, K1 uint32 }

// Philox4x32_10 is the 10-round Philox-4x32 bijection (Salmon et al., SC'11).
func Philox4x32_10(c0, c1, c2, c3, k0, k1 uint32) PxBlock {
	for r := 0; r < 10; r++ {
		p0 := uint64(0xD2511F53) * uint64(c0)
		p1 := uint64(0xCD9E8D57) * uint64(c2)
		n0 := uint32(p1>>32) ^ c1 ^ k0
		n1 := uint32(p1)
		n2 := uint32(p0>>32) ^ c3 ^ k1
		n3 := uint32(p0)
		c0, c1, c2, c3 = n0, n1, n2, n3
		k0 += 0x9E3779B9
		k1 += 0xBB67AE85
	}
	return PxBlock{c0, c1, c2, c3}
}

// Stream tags and the sub-streams of a molecule's fragmentation stream (top byte of c3).
const (
	PxTagTarget = 1
	PxTagFrag   = 2
	PxTagPcr    = 3
	PxTagSelect = 4
	PxTagStrand = 5

	PxSubHeader   = 0
	PxSubBreaks   = 1
	PxSubInterval = 2
	PxSubJumpLen  = 3
)

// PxDeriveKey: key(tag) = first two words of Philox(ctr = (tag, 'RLSI', 'MB20', 0), key = (seed_lo, seed_hi)).
// Seeds follow main.go:72-81,129-149 (-si, -sp defaulting to -si, -ss defaulting to -sp).
func PxDeriveKey(seed int64, tag uint32) PxKey {
	b := Philox4x32_10(tag, 0x524c5349, 0x4d423230, 0, uint32(uint64(seed)&0xffffffff), uint32(uint64(seed)>>32))
	return PxKey{b[0], b[1]}
}

// PxStream = key + the three upper counter words; c0 is the block index.
//
//	target length i            : key(TARGET), (i_lo, i_hi, 0), sequential cursor
//	molecule m of transcript t : key(FRAG),   (t, m_lo, sub<<24 | m_hi)
//	species (t, start, end)    : key(PCR),    (t, start, end), one block per binomial attempt
//	length class l             : key(SELECT), (l, 0, 0), candidate j = block j
//	copy j of a species        : key(STRAND), (t, start, end), draw j
type PxStream struct {
	Key        PxKey
	C1, C2, C3 uint32
	cur        uint64 // sequential cursor in 64-bit halves
	curb       uint32 // sequential cursor in whole blocks (PCR stream)
	cachedB    uint32
	have       bool
	blk        PxBlock
}

func NewPxStream(key PxKey, c1, c2, c3 uint32) *PxStream {
	return &PxStream{Key: key, C1: c1, C2: c2, C3: c3}
}

// MolStream is the fragmentation stream of molecule mol of (global) transcript tg.
func MolStream(key PxKey, tg uint32, mol uint64, sub uint32) *PxStream {
	return NewPxStream(key, tg, uint32(mol&0xffffffff), sub<<24|uint32((mol>>32)&0xffffff))
}

func (s *PxStream) Block(b uint32) PxBlock {
	return Philox4x32_10(b, s.C1, s.C2, s.C3, s.Key.K0, s.Key.K1)
}

func (b PxBlock) Lo64() uint64 { return uint64(b[0]) | uint64(b[1])<<32 }
func (b PxBlock) Hi64() uint64 { return uint64(b[2]) | uint64(b[3])<<32 }

// Draw64 is 64-bit draw number d (random access): block d>>1, half d&1.
func (s *PxStream) Draw64(d uint64) uint64 {
	b := uint32(d >> 1)
	if !s.have || s.cachedB != b {
		s.blk, s.cachedB, s.have = s.Block(b), b, true
	}
	if d&1 == 1 {
		return s.blk.Hi64()
	}
	return s.blk.Lo64()
}

// Next64 advances the sequential cursor by one 64-bit draw.
func (s *PxStream) Next64() uint64 { d := s.cur; s.cur++; return s.Draw64(d) }

// NextBlock advances the block cursor (one whole block per binomial attempt).
func (s *PxStream) NextBlock() PxBlock { b := s.curb; s.curb++; return s.Block(b) }

// PxU01: U in [0,1) from the top 53 bits.
func PxU01(x uint64) float64 { return float64(x>>11) * 1.1102230246251565404e-16 }

// PxU01Open: U in (0,1) from the top 52 bits + 1/2 (exactly representable).
func PxU01Open(x uint64) float64 { return (float64(x>>12) + 0.5) * 2.2204460492503130808e-16 }

// PxBelowBlock: bounded integer from one whole block, floor(X * n / 2^128) with X = lo64 * 2^64 + hi64
// (bias below 2^-64, no rejection loop, random access).
func PxBelowBlock(b PxBlock, n uint64) uint64 {
	x0, x1 := b.Lo64(), b.Hi64()
	loHi, _ := bits.Mul64(x1, n)    // floor(x1*n / 2^64)
	pHi, pLo := bits.Mul64(x0, n)   // x0*n
	_, carry := bits.Add64(pLo, loHi, 0)
	return pHi + carry
}

// Below draws from block `block` of the stream.
func (s *PxStream) Below(block uint32, n uint64) uint64 { return PxBelowBlock(s.Block(block), n) }
