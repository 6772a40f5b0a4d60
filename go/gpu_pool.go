// gpu_pool.go - cgo shim that plugs librlsim_gpu.so (include/rlsim_gpu.h) into sbotond/rlsim.
//
// Drop this file into the reference's src/ directory (package main).  It adds two types:
//
//	GpuPool    implements Pooler  (src/pool.go:36-48, all 11 methods)
//	GpuSampler implements Sampler (src/sampler.go:39-42, both methods)
//
// and main.go selects them when -gpu is given (INTEGRATION.md section 3).  Everything above the two batch methods
// Pooler.InitTranscripts (src/pool.go:83-135) and Sampler.SampleFragments (src/sampler.go:108-154) - args.go,
// input.go, fasta.go, report.go, fragstats.go, logging.go - stays as it is.
//
// The per-transcript methods of Pooler (AddTranscript, RegisterFragments, SampleTranscript, ...) are the reference's
// internal plumbing between Pool, Transcript and LenSampler; on the GPU path that plumbing lives on the device, so they
// end the run with the reference's own error policy (L.Fatalf, src/logging.go:67-73) if anything ever calls them.
//
// No Go toolchain exists in the image this repository is built in: the file is checked by tests/test_host.py
// (interface coverage against the reference's method lists, balanced syntax) and mirrored call for call by
// rlsim_b200/pool.py, which runs in every GPU test.
//
// Build: CGO_CFLAGS="-I<repo>/include" CGO_LDFLAGS="-L<repo>/rlsim_b200 -lrlsim_gpu -Wl,-rpath,<repo>/rlsim_b200" go build
package main

/*
#include <stdlib.h>
#include "rlsim_gpu.h"
*/
import "C"

import (
	"bufio"
	"fmt"
	"os"
	"sort"
	"unsafe"
)

// GpuPool implements Pooler on one GPU.
type GpuPool struct {
	h      *C.rlsim_handle
	gobDir string
	names  []string // transcript names in input order (Frag.String needs them)
	nrTr   int      // transcripts with a non-zero expression level (pool.go:104-106)
	rng    string   // "philox": target lengths are drawn on the device; otherwise the host Targeter's draw is imposed
	Packed bool     // hand the transcripts over as 2-bit codes + a mask of the letters outside ACGT (0.375 B per base on the link)
}

// packCode maps a letter to its 2-bit code; ok is false for every letter outside ACGT (it travels as a mask bit).
func packCode(b byte) (code byte, ok bool) {
	switch b {
	case 'A':
		return 0, true
	case 'C':
		return 1, true
	case 'G':
		return 2, true
	case 'T':
		return 3, true
	}
	return 0, false
}

// appendPacked appends seq at base position n to the packed arrays of rlsim_add_transcripts_packed (include/rlsim_gpu.h):
// codes hold 4 bases per byte, low bits first; mask one bit per base, bit i&7 of byte i>>3.
func appendPacked(codes, mask []byte, n int, seq string) ([]byte, []byte) {
	for i := 0; i < len(seq); i, n = i+1, n+1 {
		if n&3 == 0 {
			codes = append(codes, 0)
		}
		if n&7 == 0 {
			mask = append(mask, 0)
		}
		c, ok := packCode(seq[i])
		codes[n>>2] |= c << (2 * uint(n&3))
		if !ok {
			mask[n>>3] |= 1 << uint(n&7)
		}
	}
	return codes, mask
}

// GpuSampler implements Sampler.
type GpuSampler struct {
	StrandBias float64
}

func gpuFatal(h *C.rlsim_handle, rc C.int) {
	if rc != 0 {
		L.Fatalf("%s", C.GoString(C.rlsim_last_error(h))) // every error is fatal (logging.go:67-73)
	}
}

func gpuUnused(method string) {
	L.Fatalf("GpuPool.%s is not used on the GPU path: fragments, pools and draws live on the device", method)
}

// NewGpuPool replaces NewPool (pool.go:66-81).
func NewGpuPool(device int, gobDir string, rng string) *GpuPool {
	p := new(GpuPool)
	p.gobDir = gobDir
	p.rng = rng
	rc := C.rlsim_create(C.int(device), &p.h)
	gpuFatal(nil, rc)
	return p
}

var gpuFragMethods = map[string]C.int32_t{ // fragmentor.go:39-60
	"after_prim": C.RLSIM_AFTER_PRIM, "after_prim_double": C.RLSIM_AFTER_PRIM_DOUBLE,
	"after_noprim": C.RLSIM_AFTER_NOPRIM, "after_noprim_double": C.RLSIM_AFTER_NOPRIM_DOUBLE,
	"pre_prim": C.RLSIM_PRE_PRIM, "prim_jump": C.RLSIM_PRIM_JUMP,
}

var gpuCompTypes = map[string]C.int32_t{"n": C.RLSIM_COMP_N, "sn": C.RLSIM_COMP_SN, "g": C.RLSIM_COMP_G} // target_mix.go:61-67

type gpuComp struct {
	c *MixComp
	w float64
}

// fillMix copies a mixture in a fixed order.  TargetMix.Components is a Go map (target_mix.go:74-76), whose iteration order
// changes from run to run; the counter-based streams of --rng=philox need one order, so components are sorted by their
// parameters (the law of the mixture does not depend on the order).
func fillMix(dst *[C.RLSIM_MAX_COMPS]C.rlsim_mixcomp, m *TargetMix) C.int32_t {
	comps := make([]gpuComp, 0, len(m.Components))
	for c, w := range m.Components {
		comps = append(comps, gpuComp{c, w})
	}
	sort.Slice(comps, func(i, j int) bool {
		a, b := comps[i], comps[j]
		if a.c.Type != b.c.Type {
			return a.c.Type < b.c.Type
		}
		if a.c.Location != b.c.Location {
			return a.c.Location < b.c.Location
		}
		if a.c.Scale != b.c.Scale {
			return a.c.Scale < b.c.Scale
		}
		if a.c.Shape != b.c.Shape {
			return a.c.Shape < b.c.Shape
		}
		if a.c.Low != b.c.Low {
			return a.c.Low < b.c.Low
		}
		if a.c.High != b.c.High {
			return a.c.High < b.c.High
		}
		return a.w < b.w
	})
	if len(comps) > C.RLSIM_MAX_COMPS {
		L.Fatalf("The GPU path supports at most %d mixture components!", int(C.RLSIM_MAX_COMPS))
	}
	for i, cw := range comps {
		dst[i]._type = gpuCompTypes[cw.c.Type]
		dst[i].weight = C.double(cw.w)
		dst[i].location = C.double(cw.c.Location)
		dst[i].scale = C.double(cw.c.Scale)
		dst[i].shape = C.double(cw.c.Shape)
		dst[i].low = C.uint64_t(cw.c.Low)
		dst[i].high = C.uint64_t(cw.c.High)
	}
	return C.int32_t(len(comps))
}

// fillParams copies what NewFragmentor / NewTechne / NewTarget / NewLenSampler receive (main.go:98-127).
func fillParams(a *CmdArgs) C.rlsim_params {
	var p C.rlsim_params
	p.kmer_len = C.int32_t(a.KmerLength)
	code, ok := gpuFragMethods[a.FragMethod.Name]
	if !ok {
		L.Fatalf("Invalid fragmentation method.")
	}
	p.frag_method = code
	p.frag_param = C.int32_t(a.FragMethod.Param)
	p.nr_cycles = C.int32_t(a.NrCycles)
	p.priming_temp = C.double(a.PrimingTemp)
	p.frag_loss_prob = C.double(a.FragLossProb)
	p.fixed_eff = C.double(a.FixedEff)
	p.strand_bias = C.double(a.StrandBias)
	if a.RawGcEffs != nil { // thermocycler.go:62-66,94-104
		p.gc_mode = 1
		effs := Techne{}.ProcessRawGcEffs(append([]float64(nil), a.RawGcEffs...), a.MinRawGcEff)
		if len(effs) != 101 {
			L.Fatalf("Raw GC efficiency table must have 101 entries, got %d", len(effs))
		}
		for i, e := range effs {
			p.gc_raw[i] = C.double(e)
		}
	} else {
		p.gc_shape = C.double(a.GcEffParam.Shape)
		p.gc_min = C.double(a.GcEffParam.Min)
		p.gc_max = C.double(a.GcEffParam.Max)
	}
	p.len_shape = C.double(a.LenEffParam.Shape)
	p.len_min = C.double(a.LenEffParam.Min)
	p.len_max = C.double(a.LenEffParam.Max)
	if a.RawLenProbs == nil {
		p.n_target = fillMix(&p.target, a.TargetMix)
	}
	p.n_polya = fillMix(&p.polya, a.PolyAParam)
	p.seed_init = C.int64_t(a.InitSeed)
	p.seed_pcr = C.int64_t(a.PcrSeed)
	p.seed_sampling = C.int64_t(a.SamplingSeed)
	return p
}

// Configure replaces NewTarget + NewFragmentor + NewTechne (main.go:98-123): the model goes down once; the target histogram
// is either drawn on the device (--rng=philox) or imposed from the host's own Targeter; the length-efficiency table comes
// from Go's own math.Pow so that the device multiplies exactly what the report prints.
func (p *GpuPool) Configure(a *CmdArgs, tg Targeter, tc *Techne) {
	params := fillParams(a)
	gpuFatal(p.h, C.rlsim_set_model(p.h, &params))
	if a.RawLenProbs != nil {
		l, pr := a.RawLenProbs.Length, a.RawLenProbs.Prob
		gpuFatal(p.h, C.rlsim_set_raw_target(p.h, (*C.uint32_t)(unsafe.Pointer(&l[0])), (*C.double)(unsafe.Pointer(&pr[0])), C.uint32_t(len(l))))
	}
	if p.rng == "philox" {
		gpuFatal(p.h, C.rlsim_sample_target(p.h, C.int64_t(a.ReqFrags)))
	} else { // keep the host Targeter's draw (target.go:102-132)
		var ls []uint32
		var cs []uint64
		for {
			l, c, ok := tg.NextLenCount()
			if !ok {
				break
			}
			ls, cs = append(ls, l), append(cs, c)
		}
		if len(ls) > 0 {
			gpuFatal(p.h, C.rlsim_set_target_hist(p.h, (*C.uint32_t)(unsafe.Pointer(&ls[0])), (*C.uint64_t)(unsafe.Pointer(&cs[0])), C.uint32_t(len(ls))))
		}
	}
	if a.FixedEff == 0.0 && a.LenEffParam.Shape != 0.0 {
		low, high := uint32(tg.GetLow()), uint32(tg.GetHigh())
		tab := make([]float64, high-low+1)
		for i := range tab {
			tab[i] = tc.CalcLengthEff(low + uint32(i)) // thermocycler.go:149-156
		}
		gpuFatal(p.h, C.rlsim_set_len_eff(p.h, (*C.double)(unsafe.Pointer(&tab[0])), C.uint64_t(low), C.uint64_t(high)))
	}
}

// pullHist reads one per-length histogram and feeds the non-zero bins to update.
func (p *GpuPool) pullHist(kind C.int, n uint32, update func(l uint32, c uint64)) {
	if n == 0 {
		return
	}
	buf := make([]uint64, n)
	gpuFatal(p.h, C.rlsim_get_hist(p.h, kind, (*C.uint64_t)(unsafe.Pointer(&buf[0])), C.uint32_t(n)))
	for l, c := range buf {
		if c != 0 {
			update(uint32(l), c)
		}
	}
}

// InitTranscripts implements Pooler (pool.go:83-135).  The `chan *Transcript` loop of input.go:103-124 becomes batched
// rlsim_add_transcripts calls; the library copies each batch during the call (cgo pointer rules) - pageable Go memory is
// staged through the library's pinned ring - and fragments / amplifies it while the next batch is being read.
func (p *GpuPool) InitTranscripts(input Inputer, tg Targeter, fg Fragmentor, tc Thermocycler, st FragStater, GCFreq int, polyAParam *TargetMix, exprMul float64, rand Rander, pcrRand Rander) {
	L.PrintfV("Poly(A) tail distribution components:")
	for _, l := range StringToSlice(polyAParam.String()) {
		L.PrintfV("%s\n", l)
	}
	L.PrintfV("Fragmenting transcripts and amplifying fragments:\n")
	const batchBases = 256 << 20
	var bases, codes, mask []byte
	nBases := 0
	offs := []uint64{0}
	var expr []uint64
	flush := func() {
		if len(expr) == 0 {
			return
		}
		op, ep := (*C.uint64_t)(unsafe.Pointer(&offs[0])), (*C.uint64_t)(unsafe.Pointer(&expr[0]))
		if p.Packed {
			var cp, mp *C.uint8_t
			if len(codes) > 0 {
				cp, mp = (*C.uint8_t)(unsafe.Pointer(&codes[0])), (*C.uint8_t)(unsafe.Pointer(&mask[0]))
			}
			gpuFatal(p.h, C.rlsim_add_transcripts_packed(p.h, cp, mp, op, ep, C.uint32_t(len(expr))))
		} else {
			var bp *C.uint8_t
			if len(bases) > 0 {
				bp = (*C.uint8_t)(unsafe.Pointer(&bases[0]))
			}
			gpuFatal(p.h, C.rlsim_add_transcripts(p.h, bp, op, ep, C.uint32_t(len(expr))))
		}
		bases, codes, mask, nBases, offs, expr = bases[:0], codes[:0], mask[:0], 0, offs[:1], expr[:0]
	}
	in := input.(Input)
	for seq := input.NextSeq(); seq != nil; seq = input.NextSeq() { // input.go:106-122 without NewTranscript
		name, level, ok := in.ParseSeqName(seq.Name)
		if !ok {
			continue
		}
		level = uint64(float64(level) * exprMul)
		st.UpdateTrLengths(uint32(len(seq.Seq)), level)
		st.UpdateExprLevels(uint32(level))
		if level > 0 {
			p.nrTr++
		}
		p.names = append(p.names, name)
		if p.Packed {
			codes, mask = appendPacked(codes, mask, nBases, seq.Seq)
		} else {
			bases = append(bases, seq.Seq...)
		}
		nBases += len(seq.Seq)
		offs = append(offs, uint64(nBases))
		expr = append(expr, level)
		if nBases >= batchBases {
			flush()
		}
	}
	flush()
	gpuFatal(p.h, C.rlsim_build_library(p.h))
	L.PrintfV("Initialized %d transcripts.\n", p.nrTr)
	var s C.rlsim_stats
	gpuFatal(p.h, C.rlsim_get_stats(p.h, &s))
	fs := st.(*FragStats) // the histograms the report needs (fragstats.go:45-56)
	p.pullHist(C.RLSIM_H_AFTER_FRAG, uint32(s.high)+1, func(l uint32, c uint64) { fs.AfterFrag[l] += c })
	p.pullHist(C.RLSIM_H_POLYA, uint32(s.polya_max)+1, func(l uint32, c uint64) { fs.PolyALen[l] += c })
	p.pullHist(C.RLSIM_H_AFTER_PCR, uint32(s.high)+1, func(l uint32, c uint64) { st.UpdateAfterPcr(l, c) })
}

// AddTranscript implements Pooler (pool.go:137-139); transcripts reach the device through InitTranscripts.
func (p *GpuPool) AddTranscript(tr Transcripter) { gpuUnused("AddTranscript") }

// GetGobDir implements Pooler (pool.go:141-143).  Nothing is spilled: pools are device resident.
func (p *GpuPool) GetGobDir() string { return p.gobDir }

// RegisterFragments implements Pooler (pool.go:156-168); the per-length pools are built by rlsim_build_library.
func (p *GpuPool) RegisterFragments(tr Transcripter, length uint32, count uint64) {
	gpuUnused("RegisterFragments")
}

// GetTranscripts implements Pooler (pool.go:145-147).
func (p *GpuPool) GetTranscripts() []Transcripter {
	gpuUnused("GetTranscripts")
	return nil
}

// GetNrTranscripts implements Pooler (pool.go:149-151).
func (p *GpuPool) GetNrTranscripts() int { return p.nrTr }

// Flatten implements Pooler (pool.go:170-189): the flattened per-length view already exists on the device.
func (p *GpuPool) Flatten() {}

// SampleTranscript implements Pooler (pool.go:191-204); the draw runs on the device inside rlsim_sample.
func (p *GpuPool) SampleTranscript(length uint32, rand Rander) (Transcripter, bool) {
	gpuUnused("SampleTranscript")
	return nil, false
}

// String implements Pooler (pool.go:223-230).
func (p *GpuPool) String() string {
	var s C.rlsim_stats
	C.rlsim_get_stats(p.h, &s)
	return fmt.Sprintf("[ GPU pool: %d transcripts, %d species, %d fragments ]", uint64(s.n_transcripts), uint64(s.n_species), uint64(s.pool_total))
}

// JettisonLenTrCounts implements Pooler (pool.go:206-208): nothing to release per length.
func (p *GpuPool) JettisonLenTrCounts(l uint32) {}

// Cleanup implements Pooler (pool.go:210-221).
func (p *GpuPool) Cleanup() {
	C.rlsim_destroy(p.h)
	p.h = nil
}

// SampleFragments implements Sampler (sampler.go:108-154): size selection + final draw on the device, the report
// histograms of sampler.go:95-101 from the library's counters, and the records printed as FASTA (sampler.go:204-207,
// frag.go:71-126).  The text is formatted on the device and streamed to stdout.
func (sl GpuSampler) SampleFragments(pl Pooler, tg Targeter, st FragStater, rand Rander) {
	p := pl.(*GpuPool)
	gpuFatal(p.h, C.rlsim_sample(p.h, nil, nil, 0))
	var s C.rlsim_stats
	gpuFatal(p.h, C.rlsim_get_stats(p.h, &s))
	nb := uint32(s.high) + 2
	target := make([]uint64, nb)
	gpuFatal(p.h, C.rlsim_get_hist(p.h, C.RLSIM_H_TARGET, (*C.uint64_t)(unsafe.Pointer(&target[0])), C.uint32_t(nb)))
	sampled := make([]uint64, nb)
	gpuFatal(p.h, C.rlsim_get_hist(p.h, C.RLSIM_H_SAMPLED, (*C.uint64_t)(unsafe.Pointer(&sampled[0])), C.uint32_t(nb)))
	missing := make([]uint64, nb)
	gpuFatal(p.h, C.rlsim_get_hist(p.h, C.RLSIM_H_MISSING, (*C.uint64_t)(unsafe.Pointer(&missing[0])), C.uint32_t(nb)))
	for l := uint32(0); l < nb; l++ {
		if target[l] != 0 { // one entry per requested length, zero counts included (sampler.go:95,101)
			st.UpdateSampled(l, sampled[l])
			st.UpdateMissing(l, missing[l])
		}
	}
	p.pullHist(C.RLSIM_H_AFTER_SAMPLING, nb, func(l uint32, c uint64) { st.UpdateAfterSampling(l, c) })
	// names -> one buffer + offsets; the device formats every record in one pass
	var nbuf []byte
	noff := make([]uint64, 1, len(p.names)+1)
	for _, n := range p.names {
		nbuf = append(nbuf, n...)
		noff = append(noff, uint64(len(nbuf)))
	}
	if len(nbuf) == 0 {
		nbuf = []byte{0}
	}
	var size C.uint64_t
	gpuFatal(p.h, C.rlsim_format_fasta(p.h, (*C.uint8_t)(unsafe.Pointer(&nbuf[0])), (*C.uint64_t)(unsafe.Pointer(&noff[0])), nil, 0, &size))
	if size > 0 {
		out := make([]byte, uint64(size))
		gpuFatal(p.h, C.rlsim_format_fasta(p.h, (*C.uint8_t)(unsafe.Pointer(&nbuf[0])), (*C.uint64_t)(unsafe.Pointer(&noff[0])), (*C.uint8_t)(unsafe.Pointer(&out[0])), size, &size))
		w := bufio.NewWriterSize(os.Stdout, 1<<20)
		w.Write(out) // sampler.go:206
		w.Flush()
	}
	fragCount := uint64(s.n_sampled)
	st.LogSamplingRatio(fragCount)
	st.UpdateNrSampled(fragCount)
	L.PrintfV("Sampled %d fragments.\n", fragCount)
	L.PrintfV("Missing fragments: %d\n", uint64(tg.GetReqFrags())-fragCount)
}

// String implements Sampler (sampler.go:229-231).
func (sl GpuSampler) String() string {
	return "[ GPU length sampler ]"
}

// Compile-time checks: the shim satisfies the reference's interfaces.
var _ Pooler = (*GpuPool)(nil)
var _ Sampler = GpuSampler{}
