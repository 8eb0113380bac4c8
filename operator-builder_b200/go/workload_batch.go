//go:build obm_gpu
// +build obm_gpu

// Batched form of the per-manifest loop in internal/workload/v1/kinds/workload.go (same package; with the obm_gpu tag
// processManifests calls processMarkersBatched instead of processMarkers inside its loop, workload.go:224-228).
//
// The reference inspects one manifest per iteration (processMarkers -> markers.InspectForYAML, workload.go:293-297),
// i.e. one lexer per YAML node of one manifest.  Here every manifest of the workload is collected first and ALL their
// comment strings go to the GPU in one obm_lex_batch call; what happens to each manifest afterwards (re-marshalling
// the nodes, processMarkerResults, the collection rewrite of the "!!var" tags) is the reference's code, unchanged.
package kinds

import (
	"bytes"
	"fmt"
	"strings"

	"gopkg.in/yaml.v3"

	"github.com/vmware-tanzu-labs/operator-builder/internal/markers/inspect"
	"github.com/vmware-tanzu-labs/operator-builder/internal/markers/lexer"
	"github.com/vmware-tanzu-labs/operator-builder/internal/workload/v1/manifests"
	"github.com/vmware-tanzu-labs/operator-builder/internal/workload/v1/markers"
)

// inspectManifestsBatched replaces the N calls of markers.InspectForYAML (internal/workload/v1/markers/markers.go:76-88):
// pass 1 over every manifest, one GPU call, pass 2 per manifest.  markers.NewInspector(markerTypes...) is that file's
// initializeMarkerInspector (:92-115) exported together with the transform InspectForYAML passes on (transformYAML, :82):
//
//	func NewInspector(markerTypes ...MarkerType) (*inspect.Inspector, []inspect.YAMLTransformer, error) {
//		insp, err := initializeMarkerInspector(markerTypes...)
//		return insp, []inspect.YAMLTransformer{transformYAML}, err
//	}
func inspectManifestsBatched(files []*manifests.Manifest, markerTypes ...markers.MarkerType) ([][]*yaml.Node, [][]*inspect.YAMLResult, error) {
	insp, transforms, err := markers.NewInspector(markerTypes...)
	if err != nil {
		return nil, nil, fmt.Errorf("%w; error initializing markers %v", err, markerTypes)
	}

	collected := make([]*inspect.Collected, len(files))
	first := make([]int, len(files))

	var inputs [][]byte

	for i, f := range files {
		c, err := insp.CollectYAML(f.Content)
		if err != nil {
			return nil, nil, processManifestError(err, f)
		}

		collected[i], first[i] = c, len(inputs)
		inputs = append(inputs, c.Inputs...)
	}

	batch, err := lexer.LexBatch(inputs) // the whole `create api` run: one H2D, one scan, one D2H
	if err != nil {
		return nil, nil, fmt.Errorf("lexing marker comments on the GPU, %w", err)
	}

	nodes := make([][]*yaml.Node, len(files))
	results := make([][]*inspect.YAMLResult, len(files))

	for i, f := range files {
		n, r, err := insp.InspectCollected(collected[i], batch, first[i], transforms...)
		if err != nil {
			return nil, nil, processManifestError(fmt.Errorf("%w; error inspecting YAML for markers %v", err, markerTypes), f)
		}

		nodes[i], results[i] = n, r
	}

	return nodes, results, nil
}

// processMarkersBatched is processMarkers (workload.go:293-329) with the inspection result handed in.
func (ws *WorkloadSpec) processMarkersBatched(manifestFile *manifests.Manifest, nodes []*yaml.Node, markerResults []*inspect.YAMLResult,
	markerTypes ...markers.MarkerType) error {
	buf := bytes.Buffer{}

	for _, node := range nodes {
		m, err := yaml.Marshal(node)
		if err != nil {
			return processManifestError(err, manifestFile)
		}

		mustWrite(buf.WriteString("---\n"))
		mustWrite(buf.Write(m))
	}

	manifestFile.Content = buf.Bytes()

	if err := ws.processMarkerResults(markerResults); err != nil {
		return processManifestError(err, manifestFile)
	}

	if markers.ContainsMarkerType(markerTypes, markers.FieldMarkerType) &&
		markers.ContainsMarkerType(markerTypes, markers.CollectionMarkerType) {
		manifestFile.Content = []byte(strings.ReplaceAll(string(manifestFile.Content), "!!var collection", "!!var parent"))
		manifestFile.Content = []byte(strings.ReplaceAll(string(manifestFile.Content), "!!start collection", "!!start parent"))
	}

	return nil
}

// The call-site change inside processManifests (workload.go:218-228), shown in full:
//
//	ws.init()
//	nodes, results, err := inspectManifestsBatched(*ws.Manifests, markerTypes...)   // NEW: before the loop
//	if err != nil { return err }
//	uniqueNames := map[string]bool{}
//	for i, manifestFile := range *ws.Manifests {
//		err := ws.processMarkersBatched(manifestFile, nodes[i], results[i], markerTypes...) // was: ws.processMarkers(manifestFile, markerTypes...)
//		...                                                                               // the rest of the loop body is unchanged
