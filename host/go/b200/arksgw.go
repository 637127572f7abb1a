// Package b200 is the cgo binding of libarksgw.so (include/arks_gateway.h) for the arks gateway.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain (SURVEY.md §0 F5). The file is the
// reference-side binding a maintainer adds under pkg/gateway/b200/ ; it needs CGO_ENABLED=1 and a CUDA runtime base
// image instead of distroless/static (dockerfiles/Dockerfile.gateway:23,27 of the reference).
//
// It registers one implementation for the three seams selected in cmd/gateway/main.go:233-270:
//
//	-ratelimiter.type=b200  -> ratelimiter.RateLimterInterface  (pkg/gateway/ratelimiter/rate_limiter.go:21-28)
//	-quota.type=b200        -> quota.QuotaService               (pkg/gateway/quota/types.go:24-28)
//	-provider.type=b200     -> qosconfig.ConfigProvider         (pkg/gateway/qosconfig/provider.go:29-37)
//
// but the hot path does not go through those per-request interfaces any more: Server.Process enqueues the
// request/response body into the per-GPU Batcher (batcher.go) and blocks on its reply channel.
package b200

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../arks_b200 -larksgw -Wl,-rpath,${SRCDIR}/../../../arks_b200
#include <stdlib.h>
#include "arks_gateway.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Ctx owns one GPU. All calls must come from the goroutine that created it (it is locked to its OS thread:
// cgo calls block an M, and the CUDA context is bound to the thread).
type Ctx struct {
	h *C.arks_ctx
}

func NewCtx(device int, maxBatch uint32, maxBatchBytes uint64) (*Ctx, error) {
	runtime.LockOSThread()
	var h *C.arks_ctx
	if rc := C.arks_create(C.int(device), C.uint32_t(maxBatch), C.uint64_t(maxBatchBytes), &h); rc != 0 {
		return nil, fmt.Errorf("arks_create: %d (no CPU fallback)", int(rc))
	}
	return &Ctx{h: h}, nil
}

func (c *Ctx) Close() { C.arks_destroy(c.h) }

func (c *Ctx) err(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return fmt.Errorf("arksgw %d: %s", int(rc), C.GoString(C.arks_last_error(c.h)))
}

// RequestBatch mirrors arks_request_batch. The slices live in pinned rings owned by the Batcher; nothing is copied
// on the Go side and the library never calls back into Go.
type RequestBatch struct {
	N         uint32
	Bodies    []byte   // 16-byte aligned spans
	BodyOff   []uint32 // N
	BodyLen   []uint32 // N
	Tokens    []byte
	TokenOff  []uint32 // N+1
	PickRand  []uint64 // N or nil
	NowUnix   int64
}

type RequestResult struct {
	Reason   []uint8 // enum arks_reason
	Detail   []uint8
	Flags    []uint8
	Qos      []int32
	Token    []int32
	Pick     []int32
	CurUsage []int64
	LimitMax []int64
}

// SubmitRequests == HandleRequestBody for N streams (pkg/gateway/handle_request.go:83-249).
func (c *Ctx) SubmitRequests(b *RequestBatch, r *RequestResult) error {
	cb := C.arks_request_batch{
		n:            C.uint32_t(b.N),
		bodies:       (*C.uint8_t)(unsafe.Pointer(&b.Bodies[0])),
		body_off:     (*C.uint32_t)(unsafe.Pointer(&b.BodyOff[0])),
		body_len:     (*C.uint32_t)(unsafe.Pointer(&b.BodyLen[0])),
		bodies_bytes: C.uint64_t(len(b.Bodies)),
		tokens:       (*C.uint8_t)(unsafe.Pointer(&b.Tokens[0])),
		token_off:    (*C.uint32_t)(unsafe.Pointer(&b.TokenOff[0])),
		now_unix:     C.int64_t(b.NowUnix),
	}
	if b.PickRand != nil {
		cb.pick_rand = (*C.uint64_t)(unsafe.Pointer(&b.PickRand[0]))
	}
	cr := C.arks_request_result{
		reason:    (*C.uint8_t)(unsafe.Pointer(&r.Reason[0])),
		detail:    (*C.uint8_t)(unsafe.Pointer(&r.Detail[0])),
		flags:     (*C.uint8_t)(unsafe.Pointer(&r.Flags[0])),
		qos:       (*C.int32_t)(unsafe.Pointer(&r.Qos[0])),
		token:     (*C.int32_t)(unsafe.Pointer(&r.Token[0])),
		pick:      (*C.int32_t)(unsafe.Pointer(&r.Pick[0])),
		cur_usage: (*C.int64_t)(unsafe.Pointer(&r.CurUsage[0])),
		limit_max: (*C.int64_t)(unsafe.Pointer(&r.LimitMax[0])),
	}
	return c.err(C.arks_submit_request_batch(c.h, &cb, &cr))
}

type ResponseBatch struct {
	N       uint32
	Bodies  []byte
	BodyOff []uint32
	BodyLen []uint32
	Qos     []int32 // from RequestResult.Qos, carried in the stream's state
	Flags   []uint8 // ARKS_RESP_STREAM / ARKS_RESP_END_OF_STREAM
	NowUnix int64
}

type ResponseResult struct {
	Reason  []uint8
	Counted []uint8
	Usage   []int64 // 3N: prompt, completion, total
}

// SubmitResponses == HandleResponseBody (status 200) for N chunks / complete bodies
// (pkg/gateway/handle_response.go:80-268).
func (c *Ctx) SubmitResponses(b *ResponseBatch, r *ResponseResult) error {
	cb := C.arks_response_batch{
		n:            C.uint32_t(b.N),
		bodies:       (*C.uint8_t)(unsafe.Pointer(&b.Bodies[0])),
		body_off:     (*C.uint32_t)(unsafe.Pointer(&b.BodyOff[0])),
		body_len:     (*C.uint32_t)(unsafe.Pointer(&b.BodyLen[0])),
		bodies_bytes: C.uint64_t(len(b.Bodies)),
		qos:          (*C.int32_t)(unsafe.Pointer(&b.Qos[0])),
		flags:        (*C.uint8_t)(unsafe.Pointer(&b.Flags[0])),
		now_unix:     C.int64_t(b.NowUnix),
	}
	cr := C.arks_response_result{
		reason:  (*C.uint8_t)(unsafe.Pointer(&r.Reason[0])),
		counted: (*C.uint8_t)(unsafe.Pointer(&r.Counted[0])),
		usage:   (*C.int64_t)(unsafe.Pointer(&r.Usage[0])),
	}
	return c.err(C.arks_submit_response_batch(c.h, &cb, &cr))
}

// SnapshotQuota feeds the 10 s ArksQuota.status sync (pkg/gateway/qosconfig/arks_impl.go:217-300).
func (c *Ctx) SnapshotQuota(out []int64) error {
	return c.err(C.arks_snapshot_quota(c.h, (*C.int64_t)(unsafe.Pointer(&out[0]))))
}

// SetQuotaUsage == quota.QuotaService.SetUsage; used at start-up to restore usage from ArksQuota.status
// (done correctly here; the reference writes 0, arks_impl.go:263-267).
func (c *Ctx) SetQuotaUsage(quota uint32, usage [3]int64) error {
	return c.err(C.arks_set_quota_usage(c.h, C.uint32_t(quota), (*C.int64_t)(unsafe.Pointer(&usage[0]))))
}

// SyncQuotaUsage runs the body of syncQuotaUsage (qosconfig/arks_impl.go:217-300) for every ArksQuota in one call.
// present[q] / used[3q..3q+2] carry Status.QuotaStatus in and out; action[q]&1: write the status back to the CR.
func (c *Ctx) SyncQuotaUsage(restore bool, present []uint32, used []int64, action []uint8) error {
	mode := C.int(C.ARKS_SYNC_REFERENCE)
	if restore {
		mode = C.int(C.ARKS_SYNC_RESTORE)
	}
	return c.err(C.arks_sync_quota_usage(c.h, mode, (*C.uint32_t)(unsafe.Pointer(&present[0])),
		(*C.int64_t)(unsafe.Pointer(&used[0])), (*C.uint8_t)(unsafe.Pointer(&action[0]))))
}

// EnableMetrics / SnapshotMetrics: the series of pkg/gateway/metrics that are functions of the request stream, kept on
// the device (ARKS_METRIC_COLS int64 per qos entry); the Prometheus collector reads them at scrape time.
func (c *Ctx) EnableMetrics(on bool) error {
	v := C.int(0)
	if on {
		v = 1
	}
	return c.err(C.arks_enable_metrics(c.h, v))
}
func (c *Ctx) SnapshotMetrics(rows []int64) error {
	return c.err(C.arks_snapshot_metrics(c.h, (*C.int64_t)(unsafe.Pointer(&rows[0]))))
}

// AllocPinned returns page-locked memory for batch staging as a byte slice (freed with FreePinned).
func AllocPinned(n int) []byte {
	p := C.arks_alloc_pinned(C.size_t(n))
	if p == nil {
		return nil
	}
	return unsafe.Slice((*byte)(p), n)
}
func FreePinned(b []byte) { C.arks_free_pinned(unsafe.Pointer(&b[0])) }
