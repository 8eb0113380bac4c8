//go:build obm_gpu
// +build obm_gpu

// Batched form of internal/markers/inspect/yaml.go (same package; with the obm_gpu tag this file replaces yaml.go).
//
// The reference runs one parser -- hence one lexer goroutine with a 4 KiB bufio.Reader -- per YAML node:
// inspectYAMLComments calls s.parse(Head + "\n" + Line + "\n" + Foot) (inspect/yaml.go:89-95, inspector.go:21-25).
// Here the same walk is done twice:
//
//	pass 1  visit the nodes in exactly the order inspectYAML / inspectYAMLMap / inspectYAMLComments visit them
//	        (yaml.go:62-107) and only COLLECT the comment strings
//	gpu     one lexer.LexBatch over all of them (one obm_lex_batch call; CollectYAML / InspectCollected below let
//	        internal/workload collect over EVERY manifest of a `create api` run before that single call)
//	pass 2  visit again in the same order; visit k takes batch.Lexer(k) and runs the UNCHANGED parser on it
//
// The lexeme stream is bit-identical to the pure-Go lexer's, parser/ and marker/ are untouched, so the Results --
// MarkerText included -- and everything scaffolded from them are identical by construction.
package inspect

import (
	"bytes"
	"errors"
	"fmt"
	"io"

	"gopkg.in/yaml.v3"

	"github.com/vmware-tanzu-labs/operator-builder/internal/markers/lexer"
	"github.com/vmware-tanzu-labs/operator-builder/internal/markers/parser"
)

type YAMLResult struct {
	*parser.Result
	Nodes []*yaml.Node
}

// Collected is pass 1's output for one manifest: its decoded documents and the comment strings of every visit.
type Collected struct {
	Nodes  []*yaml.Node
	Inputs [][]byte
}

// visit is the shared traversal.  f is called once per inspectYAMLComments node (yaml.go:92-95), in order.
func visitYAML(nodes []*yaml.Node, f func(group []*yaml.Node, node *yaml.Node)) {
	for _, node := range nodes {
		f([]*yaml.Node{node}, node)

		if node.Kind == yaml.MappingNode {
			visitYAMLMap(node.Content, f)
		} else if node.Content != nil {
			visitYAML(node.Content, f)
		}
	}
}

func visitYAMLMap(nodes []*yaml.Node, f func(group []*yaml.Node, node *yaml.Node)) {
	for i := 0; i < len(nodes); i += 2 {
		group := []*yaml.Node{nodes[i], nodes[i+1]}
		f(group, nodes[i])
		f(group, nodes[i+1])

		if nodes[i+1].Kind == yaml.MappingNode {
			visitYAMLMap(nodes[i+1].Content, f)
		} else {
			visitYAML(nodes[i+1].Content, f)
		}
	}
}

// CollectYAML is pass 1 for one manifest (the decode loop is yaml.go:25-37).
func (s *Inspector) CollectYAML(data []byte) (*Collected, error) {
	c := &Collected{}
	dec := yaml.NewDecoder(bytes.NewReader(data))

	for {
		var node yaml.Node

		if err := dec.Decode(&node); errors.Is(err, io.EOF) {
			break
		} else if err != nil {
			return nil, fmt.Errorf("error unmarshaling yaml, %w", err)
		}

		c.Nodes = append(c.Nodes, &node)
	}

	visitYAML(c.Nodes, func(_ []*yaml.Node, node *yaml.Node) {
		c.Inputs = append(c.Inputs, []byte(fmt.Sprintf("%s\n%s\n%s", node.HeadComment, node.LineComment, node.FootComment)))
	})

	return c, nil
}

// InspectCollected is pass 2 for one manifest: batch holds the lexed inputs of one or more manifests, first is the
// index of this manifest's first input inside it.  Returns what InspectYAML returns.
func (s *Inspector) InspectCollected(c *Collected, batch *lexer.Batch, first int, transforms ...YAMLTransformer) ([]*yaml.Node, []*YAMLResult, error) {
	var results []*YAMLResult

	k := first
	var pending []*parser.Result // results of the visits of the current inspectYAMLComments call
	var pendingGroup []*yaml.Node
	flush := func() {
		for _, marker := range pending {
			results = append(results, &YAMLResult{Result: marker, Nodes: pendingGroup})
		}
		pending, pendingGroup = nil, nil
	}

	visitYAML(c.Nodes, func(group []*yaml.Node, _ *yaml.Node) {
		// inspectYAMLComments(nodes...) parses every node of the group, then wraps all of its markers with the whole
		// group (yaml.go:92-105); a group is one node (yaml.go:64) or a key/value pair (yaml.go:78)
		if pendingGroup != nil && (len(group) != len(pendingGroup) || group[0] != pendingGroup[0]) {
			flush()
		}
		pendingGroup = group
		p := parser.NewParserFromLexer(batch.Lexer(k), s.Registry)
		pending = append(pending, p.Parse()...)
		k++
	})
	flush()

	for _, result := range results {
		if v, ok := result.Result.Object.(error); ok {
			return c.Nodes, results, v
		}
	}

	for _, transform := range transforms {
		if err := transform(results...); err != nil {
			return c.Nodes, nil, err
		}
	}

	return c.Nodes, results, nil
}

// InspectYAML keeps the reference's signature (yaml.go:22): one manifest, one GPU call.
func (s *Inspector) InspectYAML(data []byte, transforms ...YAMLTransformer) ([]*yaml.Node, []*YAMLResult, error) {
	c, err := s.CollectYAML(data)
	if err != nil {
		return nil, nil, err
	}

	batch, err := lexer.LexBatch(c.Inputs)
	if err != nil {
		return nil, nil, fmt.Errorf("lexing marker comments on the GPU, %w", err)
	}

	return s.InspectCollected(c, batch, 0, transforms...)
}
