//go:build obm_gpu
// +build obm_gpu

// Addition to internal/markers/parser (same package, one new constructor).
//
// NewParser (parser/parser.go:28-50) builds its own lexer from the input string and starts it in a goroutine
// (`lexer.NewLexer(bytes.NewBufferString(input))` at :35, `go p.lexer.Run()` at :47).  The batched inspector has
// already lexed every node's comment string on the GPU, so it hands the parser a pre-lexed stream instead; every
// other field is initialised exactly as NewParser does, and Parse / Run / the state functions are unchanged.
package parser

import "github.com/vmware-tanzu-labs/operator-builder/internal/markers/lexer"

func NewParserFromLexer(lx *lexer.Lexer, registry Registry) *Parser {
	const bufferSize = 3

	return &Parser{
		name:        "Marker Parser",
		scopeBuffer: "",
		registry:    registry,
		lexer:       lx, // already run: NextLexeme replays the GPU's tuples (no goroutine to start)
		currentLexeme: lexer.Lexeme{
			Type:  lexer.LexemeError,
			Value: "",
		},
		peekStack: [3]lexer.Lexeme{},
		peekCount: 0,
		stack:     make([]stateFn, 0),
		state:     startParse,
		items:     make(chan *Result, bufferSize),
	}
}
