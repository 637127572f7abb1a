package b200

// Micro-batcher: N ext_proc stream goroutines enqueue; one goroutine per GPU, locked to its OS thread, drains the
// queue into pinned SoA buffers and makes ONE cgo call per micro-batch (cgo calls are ~100 ns but block an M; CUDA
// submissions are ~10 us: keep them few and large). Batches are cut by deadline, not by count, so that the added
// latency stays inside the p99 budget (BASELINE.json: < 200 us): a batch is submitted when it is `MaxBatch` long or
// `MaxWait` after its first element, whichever comes first.
//
// NOT COMPILED HERE (no Go toolchain in the build image). A C++ equivalent of this batching logic is exercised by
// bench.py's latency sweep through the same C ABI.

import (
	"time"
)

type reqItem struct {
	body  []byte
	token []byte
	reply chan reqReply
}

type reqReply struct {
	Reason, Detail, Flags uint8
	Qos, Token, Pick      int32
	CurUsage, LimitMax    int64
}

type Batcher struct {
	ctx      *Ctx
	in       chan reqItem
	MaxBatch int
	MaxWait  time.Duration
}

func NewBatcher(device int, maxBatch int, maxWait time.Duration) *Batcher {
	b := &Batcher{in: make(chan reqItem, 4*maxBatch), MaxBatch: maxBatch, MaxWait: maxWait}
	ready := make(chan error)
	go func() {
		ctx, err := NewCtx(device, uint32(maxBatch), uint64(maxBatch)*4096) // LockOSThread inside
		b.ctx = ctx
		ready <- err
		if err == nil {
			b.loop()
		}
	}()
	if err := <-ready; err != nil {
		panic(err)
	}
	return b
}

// HandleRequestBody is what Server.Process calls instead of s.HandleRequestBody (pkg/gateway/gateway.go:111-112).
func (b *Batcher) HandleRequestBody(body, token []byte) reqReply {
	it := reqItem{body: body, token: token, reply: make(chan reqReply, 1)}
	b.in <- it
	return <-it.reply
}

func (b *Batcher) loop() {
	batch := &RequestBatch{}
	res := &RequestResult{}
	pending := make([]reqItem, 0, b.MaxBatch)
	for first := range b.in {
		pending = append(pending[:0], first)
		deadline := time.NewTimer(b.MaxWait)
	fill:
		for len(pending) < b.MaxBatch {
			select {
			case it := <-b.in:
				pending = append(pending, it)
			case <-deadline.C:
				break fill
			}
		}
		deadline.Stop()
		packRequests(batch, res, pending, time.Now().Unix()) // bodies at 16-byte aligned offsets, SoA results sized
		if err := b.ctx.SubmitRequests(batch, res); err != nil {
			for _, it := range pending {
				it.reply <- reqReply{Reason: 255}
			}
			continue
		}
		for i, it := range pending {
			it.reply <- reqReply{res.Reason[i], res.Detail[i], res.Flags[i], res.Qos[i], res.Token[i], res.Pick[i],
				res.CurUsage[i], res.LimitMax[i]}
		}
	}
}

func packRequests(b *RequestBatch, r *RequestResult, items []reqItem, now int64) {
	n := len(items)
	b.N, b.NowUnix = uint32(n), now
	b.Bodies, b.Tokens = b.Bodies[:0], b.Tokens[:0]
	b.BodyOff, b.BodyLen, b.TokenOff = b.BodyOff[:0], b.BodyLen[:0], append(b.TokenOff[:0], 0)
	for _, it := range items {
		for len(b.Bodies)%16 != 0 {
			b.Bodies = append(b.Bodies, 0)
		}
		b.BodyOff = append(b.BodyOff, uint32(len(b.Bodies)))
		b.BodyLen = append(b.BodyLen, uint32(len(it.body)))
		b.Bodies = append(b.Bodies, it.body...)
		b.Tokens = append(b.Tokens, it.token...)
		b.TokenOff = append(b.TokenOff, uint32(len(b.Tokens)))
	}
	for len(b.Bodies)%16 != 0 {
		b.Bodies = append(b.Bodies, 0)
	}
	grow := func(n int) {
		if cap(r.Reason) < n {
			r.Reason, r.Detail, r.Flags = make([]uint8, n), make([]uint8, n), make([]uint8, n)
			r.Qos, r.Token, r.Pick = make([]int32, n), make([]int32, n), make([]int32, n)
			r.CurUsage, r.LimitMax = make([]int64, n), make([]int64, n)
		}
	}
	grow(n)
}
