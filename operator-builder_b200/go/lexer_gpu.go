//go:build obm_gpu
// +build obm_gpu

// Package lexer -- GPU-backed drop-in for internal/markers/lexer.
//
// Build with `-tags obm_gpu` (CGO_ENABLED=1, -lobmarkers); without the tag the reference's pure-Go
// lexer files are compiled instead, so release builds (CGO_ENABLED=0, .goreleaser.yml:9) are
// unchanged.  This file replaces lexer.go / state.go / peek.go / consume.go / discard.go /
// position.go's method bodies; lexeme.go (LexemeType constants, Lexeme struct) is kept as is.
//
// It keeps the three-method surface the parser consumes (parser/parser.go:35,47; parser/peek.go:19;
// parser/position.go:15,30):
//
//	func NewLexer(r io.Reader) *Lexer     lexer.go:27
//	func (l *Lexer) Run()                 lexer.go:43
//	func (l *Lexer) NextLexeme() Lexeme   lexer.go:51
//
// and adds the batch entry point the restructured inspect package uses (INTEGRATION.md):
//
//	func LexBatch(docs [][]byte) (*Batch, error)   one obm_lex_batch() for every comment string of
//	                                               every manifest of a `create api` run
//	func (b *Batch) Lexer(i int) *Lexer            the pre-lexed stream of document i
//
// NOTE: this image has no Go toolchain; the file is written against include/obmarkers.h and is
// exercised through the identical C ABI by the Python mirror (operator-builder_b200/lexer.py).
package lexer

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/.. -lobmarkers -lstdc++
#include <stdlib.h>
#include "obmarkers.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"runtime"
	"sync"
	"unsafe"
)

// position mirrors lexer/position.go:12-15 (unexported fields, printed with %+v by parser/error.go).
type position struct {
	line   int
	column int
}

// Lexer replays one document's tuples as the reference's Lexeme sequence.
type Lexer struct {
	stream *C.obm_stream
	keep   *Batch // keeps the C buffers alive
	closed bool
	failure string // infrastructure failure, reported in band by the first NextLexeme
}

// Batch owns the C copies of the documents and the tuple stream of one obm_lex_batch call.
type Batch struct {
	bytes    unsafe.Pointer // page-locked C memory: packed documents
	docOff   []C.uint64_t
	tuples   unsafe.Pointer // page-locked C memory: obm_tuple[]
	tupleOff []C.uint64_t
	Stats    C.obm_stats
}

var (
	handleOnce sync.Once
	handle     *C.obm_handle
	handleErr  error
	handleMu   sync.Mutex // the handle is single-threaded (one CUDA stream)
)

func getHandle() (*C.obm_handle, error) {
	handleOnce.Do(func() {
		var h *C.obm_handle
		if rc := C.obm_create(0, &h); rc != C.OBM_OK {
			handleErr = fmt.Errorf("obm_create: %s (no CPU fallback in the obm_gpu build)", C.GoString(C.obm_last_error(nil)))
			return
		}
		handle = h
	})
	return handle, handleErr
}

// LexBatch lexes every document in ONE obm_lex_batch call.  The packed documents and the tuple buffer live in
// page-locked C memory (obm_pinned_alloc) so that the library's H2D / D2H copies run at DMA speed; the tuple buffer is
// sized from the input (manifests need ~0.35 B of tuples per byte, 1 B/byte + 4 tuples per document is a safe first
// guess) and grown once if the library answers OBM_E_CAPACITY with the exact count.
func LexBatch(docs [][]byte) (*Batch, error) {
	h, err := getHandle()
	if err != nil {
		return nil, err
	}
	total := 0
	for _, d := range docs {
		total += len(d)
	}
	b := &Batch{docOff: make([]C.uint64_t, len(docs)+1), tupleOff: make([]C.uint64_t, len(docs)+1)}
	b.bytes = C.obm_pinned_alloc(C.uint64_t(total + 64))
	if b.bytes == nil {
		return nil, errors.New("obm_pinned_alloc failed")
	}
	runtime.SetFinalizer(b, func(b *Batch) { C.obm_pinned_free(b.bytes); C.obm_pinned_free(b.tuples) })
	off := 0
	for i, d := range docs {
		b.docOff[i] = C.uint64_t(off)
		if len(d) > 0 {
			copy(unsafe.Slice((*byte)(unsafe.Add(b.bytes, off)), len(d)), d)
		}
		off += len(d)
	}
	b.docOff[len(docs)] = C.uint64_t(off)

	handleMu.Lock()
	defer handleMu.Unlock()
	capTuples := C.uint64_t(total/8 + 4*len(docs) + 64)
	for attempt := 0; attempt < 2; attempt++ {
		if b.tuples != nil {
			C.obm_pinned_free(b.tuples)
		}
		b.tuples = C.obm_pinned_alloc(capTuples * 8)
		if b.tuples == nil {
			return nil, errors.New("obm_pinned_alloc failed")
		}
		var need C.uint64_t
		rc := C.obm_lex_batch(h, (*C.uint8_t)(b.bytes), &b.docOff[0], C.uint32_t(len(docs)), (*C.obm_tuple)(b.tuples), capTuples, &need,
			&b.tupleOff[0], &b.Stats)
		if rc == C.OBM_OK {
			return b, nil
		}
		if rc != C.OBM_E_CAPACITY {
			return nil, errors.New(C.GoString(C.obm_last_error(h)))
		}
		capTuples = need // exact, reported by the library: the second attempt cannot fail for capacity
	}
	return nil, errors.New("obm_lex_batch: capacity retry failed")
}

// Lexer returns the pre-lexed stream of document i.
func (b *Batch) Lexer(i int) *Lexer {
	doc := unsafe.Add(b.bytes, int(b.docOff[i]))
	n := b.docOff[i+1] - b.docOff[i]
	t := unsafe.Add(b.tuples, int(b.tupleOff[i])*8)
	nt := b.tupleOff[i+1] - b.tupleOff[i]
	l := &Lexer{stream: C.obm_stream_new((*C.uint8_t)(doc), n, (*C.obm_tuple)(t), nt), keep: b}
	runtime.SetFinalizer(l, func(l *Lexer) { C.obm_stream_free(l.stream) })
	return l
}

// NewLexer creates a lexer for the input reader (lexer.go:27): a batch of one document.  An infrastructure failure
// (no device, CUDA error) cannot be returned from this signature; it is surfaced the way the reference surfaces lexical
// errors -- in band: the first NextLexeme yields a LexemeError whose Value carries the library's message, then the
// stream is closed.  (The parser turns a LexemeError into an error Result, parser/state.go:33-37.)
func NewLexer(r io.Reader) *Lexer {
	data, rerr := io.ReadAll(r)
	if rerr != nil {
		return &Lexer{failure: "reading the lexer input: " + rerr.Error()}
	}
	b, err := LexBatch([][]byte{data})
	if err != nil {
		return &Lexer{failure: "operator-builder GPU lexer: " + err.Error()}
	}
	return b.Lexer(0)
}

// Run is a no-op: the scan already ran on the GPU (lexer.go:43 ran the state machine here).
func (l *Lexer) Run() {}

// NextLexeme returns the next lexeme; after the last one it returns the zero Lexeme, like a receive
// from the closed channel in the reference (lexer.go:47,51-53).
func (l *Lexer) NextLexeme() Lexeme {
	if l.failure != "" && !l.closed {
		l.closed = true
		return Lexeme{Type: LexemeError, Value: l.failure}
	}
	if l.closed || l.stream == nil {
		return Lexeme{}
	}
	var lx C.obm_lexeme
	if C.obm_stream_next(l.stream, &lx) == 0 {
		l.closed = true
		return Lexeme{}
	}
	return Lexeme{
		Type:  LexemeType(lx._type),
		Value: C.GoStringN((*C.char)(unsafe.Pointer(lx.value)), C.int(lx.value_len)),
		Pos:   position{line: int(lx.line), column: int(lx.column)},
	}
}
