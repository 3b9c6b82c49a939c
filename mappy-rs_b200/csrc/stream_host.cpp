/* stream_host.cpp -- mmg_submit / mmg_next: the streaming face of the C ABI.
 *
 * Replaces, for a host that feeds reads one by one, the work queue + worker threads + result queue of the
 * reference (/root/reference/src/lib.rs:297-309, 541-636: `enable_threading` spawns workers that pop (id, seq) from
 * a bounded queue, call mm_map and push (id, mappings) to a result queue; `map_batch` pushes, the iterator pops in
 * COMPLETION order, src/lib.rs:972-991).  Here submit() copies the reads into page-locked staging owned by the
 * library and returns; ONE worker thread per aligner hands device-sized batches to mmg_map_batch as soon as enough
 * bases are queued, the producer pauses or flushes; next() delivers one read's hits at a time, in submission order
 * within a batch and batch by batch (an allowed completion order).  Built on the public entry points only.
 */
#include <string.h>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "mmg_internal.h"

namespace {

struct Pending {                 /* reads accepted by submit(), not yet mapped */
	char *bases; size_t cap, used;      /* page-locked */
	std::vector<uint64_t> off;          /* n + 1 */
	std::vector<uint64_t> ids;
};
struct Done {                    /* a mapped batch being handed out */
	mmg_batch *b;
	std::vector<uint64_t> ids;
	uint32_t next;                      /* next read to deliver */
	int refs;                           /* results handed out and not yet released */
};
struct Stream {
	mmg_aligner *al;
	std::mutex mu;
	std::condition_variable cv_work, cv_done;
	Pending *fill;                      /* being filled by submit() */
	std::deque<Done*> done;
	std::thread worker;
	uint64_t drain_bases, queued_bases, in_flight;
	std::chrono::steady_clock::time_point last_push;
	bool flush, stop;
	int err; std::string err_msg;
	Stream() : al(0), fill(0), drain_bases((uint64_t)32 << 20), queued_bases(0), in_flight(0), flush(false), stop(false), err(0) {}
};

std::mutex g_mu;
std::map<mmg_aligner*, Stream*> g_streams;

Pending *pending_new(size_t cap)
{
	Pending *p = new Pending();
	void *m = 0;
	if (mmg_host_alloc(cap, &m) != MMG_OK) { delete p; return 0; }
	p->bases = (char*)m, p->cap = cap, p->used = 0;
	p->off.push_back(0);
	return p;
}
void pending_free(Pending *p) { if (p) { mmg_host_free(p->bases); delete p; } }

void worker_main(Stream *s)
{
	const auto idle = std::chrono::milliseconds(20);
	for (;;) {
		Pending *job = 0;
		{
			std::unique_lock<std::mutex> lk(s->mu);
			for (;;) {
				if (s->stop) return;
				const bool have = s->fill && s->fill->ids.size() > 0;
				if (have && (s->flush || s->queued_bases >= s->drain_bases || std::chrono::steady_clock::now() - s->last_push >= idle)) break;
				if (!have && s->flush) s->flush = false;
				s->cv_work.wait_for(lk, have ? idle / 2 : std::chrono::milliseconds(50));
			}
			job = s->fill, s->fill = 0, s->queued_bases = 0;
			s->in_flight += job->ids.size();
		}
		mmg_batch *b = 0;
		const int rc = mmg_map_batch(s->al, job->bases, job->off.data(), (uint32_t)job->ids.size(), &b);
		{
			std::lock_guard<std::mutex> lk(s->mu);
			if (rc != MMG_OK) s->err = rc, s->err_msg = mmg_last_error(), s->in_flight -= job->ids.size();
			else {
				Done *d = new Done();
				d->b = b, d->ids.swap(job->ids), d->next = 0, d->refs = 0;
				s->done.push_back(d);
			}
		}
		s->cv_done.notify_all();
		pending_free(job);
	}
}

Stream *stream_of(mmg_aligner *al, bool create)
{
	std::lock_guard<std::mutex> g(g_mu);
	std::map<mmg_aligner*, Stream*>::iterator it = g_streams.find(al);
	if (it != g_streams.end()) return it->second;
	if (!create) return 0;
	Stream *s = new Stream();
	s->al = al, s->last_push = std::chrono::steady_clock::now();
	s->worker = std::thread(worker_main, s);
	g_streams[al] = s;
	return s;
}

} // namespace

/* called by mmg_aligner_destroy before the device state goes away */
void mmg_stream_shutdown(mmg_aligner *al)
{
	Stream *s = 0;
	{
		std::lock_guard<std::mutex> g(g_mu);
		std::map<mmg_aligner*, Stream*>::iterator it = g_streams.find(al);
		if (it == g_streams.end()) return;
		s = it->second;
		g_streams.erase(it);
	}
	{ std::lock_guard<std::mutex> lk(s->mu); s->stop = true; }
	s->cv_work.notify_all();
	s->worker.join();
	pending_free(s->fill);
	for (size_t i = 0; i < s->done.size(); ++i) { mmg_batch_destroy(s->done[i]->b); delete s->done[i]; }
	delete s;
}

extern "C" {

int mmg_submit(mmg_aligner *al, const char *bases, const uint64_t *offsets, uint32_t n_reads, uint64_t first_id)
{
	Stream *s = stream_of(al, true);
	const uint64_t nb = n_reads ? offsets[n_reads] - offsets[0] : 0;
	std::unique_lock<std::mutex> lk(s->mu);
	if (s->err) { mmg_set_error("%s", s->err_msg.c_str()); return s->err; }
	if (n_reads == 0) return MMG_OK;
	if (!s->fill || s->fill->used + nb > s->fill->cap) {
		if (s->fill && s->fill->ids.size()) { /* the staging block is full: let the worker take it, then start a new one */
			s->flush = true;
			s->cv_work.notify_all();
			while (s->fill && !s->stop && !s->err) s->cv_done.wait_for(lk, std::chrono::milliseconds(5));
		}
		if (!s->fill) {
			size_t cap = (size_t)(s->drain_bases * 2);
			if (cap < nb) cap = nb;
			s->fill = pending_new(cap);
			if (!s->fill) return MMG_ENOMEM;
		} else if (s->fill->used + nb > s->fill->cap) { mmg_set_error("one submission of %llu bases exceeds the staging block", (unsigned long long)nb); return MMG_EINVAL; }
	}
	Pending *p = s->fill;
	memcpy(p->bases + p->used, bases + offsets[0], nb);
	for (uint32_t i = 0; i < n_reads; ++i) {
		p->off.push_back(p->used + (offsets[i + 1] - offsets[0]));
		p->ids.push_back(first_id + i);
	}
	p->used += nb, s->queued_bases += nb;
	s->last_push = std::chrono::steady_clock::now();
	if (s->queued_bases >= s->drain_bases) s->cv_work.notify_all();
	return MMG_OK;
}

int mmg_flush(mmg_aligner *al)
{
	Stream *s = stream_of(al, false);
	if (!s) return MMG_OK;
	{ std::lock_guard<std::mutex> lk(s->mu); s->flush = true; }
	s->cv_work.notify_all();
	return MMG_OK;
}

int mmg_next(mmg_aligner *al, mmg_result_t *out, int timeout_ms)
{
	memset(out, 0, sizeof(*out));
	Stream *s = stream_of(al, false);
	if (!s) return 0;
	std::unique_lock<std::mutex> lk(s->mu);
	const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms < 0 ? 0 : timeout_ms);
	for (;;) {
		while (!s->done.empty() && s->done.front()->next == s->done.front()->ids.size() && s->done.front()->refs == 0) {
			Done *d = s->done.front();       /* everything delivered and released */
			s->done.pop_front();
			mmg_batch_destroy(d->b);
			delete d;
		}
		for (size_t k = 0; k < s->done.size(); ++k) {
			Done *d = s->done[k];
			if (d->next < d->ids.size()) {
				const uint32_t i = d->next++;
				const uint64_t *ho = mmg_batch_hit_off(d->b);
				out->read_id = d->ids[i], out->n_hits = (uint32_t)(ho[i + 1] - ho[i]);
				out->hits = mmg_batch_hits(d->b) + ho[i], out->cigar = mmg_batch_cigar(d->b), out->owner = d;
				++d->refs, --s->in_flight;
				return 1;
			}
		}
		if (s->err) { mmg_set_error("%s", s->err_msg.c_str()); return s->err; }
		const bool pending = s->in_flight > 0 || (s->fill && s->fill->ids.size() > 0);
		if (!pending) return 0;                      /* nothing submitted that has not been delivered */
		if (timeout_ms >= 0 && std::chrono::steady_clock::now() >= deadline) return 0;
		if (timeout_ms < 0) s->cv_done.wait_for(lk, std::chrono::milliseconds(50));
		else s->cv_done.wait_until(lk, deadline);
	}
}

void mmg_result_release(mmg_aligner *al, mmg_result_t *r)
{
	if (!r || !r->owner) return;
	Stream *s = stream_of(al, false);
	if (s) { std::lock_guard<std::mutex> lk(s->mu); --((Done*)r->owner)->refs; }
	r->owner = 0, r->hits = 0, r->cigar = 0, r->n_hits = 0;
}

} // extern "C"
