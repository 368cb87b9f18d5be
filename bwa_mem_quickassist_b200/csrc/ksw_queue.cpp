// ksw_queue.cpp — one submission queue per GPU shared by all host threads (include/ksw_b200.h; SURVEY.md 8(f) rank 1,
// the part the rounds scheduler left open: cross-thread batch coalescing).  Built on the public C ABI only.  The queue
// owns a few contexts ("lanes", two by default) on its device and NO thread of its own: a submitter that finds a lane free becomes the
// leader of a batch — it takes everything that is pending for the same kind of work and the same scoring (its own
// submission or not), runs it as one GPU batch on that lane from its own thread, and wakes the submitters it served;
// submitters that find every lane busy sleep until they are served or a lane frees up.  No delay is added to collect
// work: while the lanes run, the next submissions pile up and become the next leader's batch.  (The first version had a
// server thread per lane; with every core busy seeding, the hand-over to a sleeping server thread cost a scheduler
// wake-up per batch, and a `bwa mem -t 16` chunk took anything between 2.2 and 3.4 s; led by the submitters
// themselves it is as steady as private contexts.)  Several lanes, so that the copies, the device-side packing and the
// per-batch host work of one merged batch overlap the kernels of the others.
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ksw_b200.h"

namespace {

struct Sub {                           // one blocked submitter
	int kind = 0;                      // 0: extension against the resident reference, 1: global alignment, 2: local alignment
	ksw_b200_cfg_t cfg;
	int64_t n = 0;
	const ksw_b200_rjob_t *rjobs = nullptr;
	const ksw_b200_gjob_t *gjobs = nullptr;
	const ksw_b200_ajob_t *ajobs = nullptr;
	ksw_b200_ares_t *ares = nullptr;
	const uint8_t *qpool = nullptr, *tpool = nullptr;
	size_t qbytes = 0, tbytes = 0;
	ksw_b200_res_t *res = nullptr;
	ksw_b200_gres_t *gres = nullptr;
	uint32_t *cigar = nullptr;         // malloc'd for the submitter
	int64_t n_cigar = 0;
	int rc = 0;
	bool done = false;
	std::condition_variable cv;
};

bool same_cfg(const ksw_b200_cfg_t &a, const ksw_b200_cfg_t &b)
{
	return memcmp(a.mat, b.mat, 25) == 0 && a.m == b.m && a.o_del == b.o_del && a.e_del == b.e_del && a.o_ins == b.o_ins &&
	       a.e_ins == b.e_ins && a.zdrop == b.zdrop && a.end_bonus == b.end_bonus;
}

} // namespace

#define KSW_QUEUE_LANES 2               /* contexts per GPU (KSW_B200_QUEUE_LANES, 1..16): the copies, the device-side packing and
                                         * the per-batch host work of one merged batch run under the kernels of the others */
struct Server {
	ksw_b200_ctx_t *ctx = nullptr;
	bool busy = false;                 // a leader is running a batch on this lane
	// merged global-alignment batch
	std::vector<ksw_b200_gjob_t> gj;
	std::vector<ksw_b200_gres_t> gr;
	std::vector<uint8_t> gq, gt;
	// merged local-alignment batch (the byte pools above are shared: a lane runs one batch at a time)
	std::vector<ksw_b200_ajob_t> aj;
	std::vector<ksw_b200_ares_t> ar;
};

struct ksw_b200_queue {
	std::vector<Server> srv;
	std::mutex mu;
	std::condition_variable cv_idle;   // destroy waits here for the lanes to drain
	std::deque<Sub *> pending;
	bool stop = false;
	std::string err;
	int64_t n_batches = 0, n_subs = 0;
};

namespace {

void run_extend(Server *q, std::vector<Sub *> &grp)
{
	std::vector<ksw_b200_rseg_t> segs(grp.size());
	for (size_t i = 0; i < grp.size(); ++i) {
		segs[i].n = grp[i]->n; segs[i].jobs = grp[i]->rjobs; segs[i].qpool = grp[i]->qpool;
		segs[i].qpool_bytes = grp[i]->qbytes; segs[i].res = grp[i]->res;
	}
	const int rc = ksw_b200_extend_batch_ref_segs(q->ctx, &grp[0]->cfg, (int)segs.size(), segs.data());
	for (Sub *s : grp) s->rc = rc;
}

void run_global(Server *q, std::vector<Sub *> &grp)
{
	// merge: job records with rebased offsets, the pools one after the other (their sizes are the jobs' extents)
	size_t total = 0, nq = 0, nt = 0;
	for (Sub *s : grp) {
		total += (size_t)s->n;
		size_t qb = 0, tb = 0;
		for (int64_t k = 0; k < s->n; ++k) {
			qb = std::max<size_t>(qb, (size_t)s->gjobs[k].q_off + (size_t)std::max(s->gjobs[k].qlen, 0));
			tb = std::max<size_t>(tb, (size_t)s->gjobs[k].t_off + (size_t)std::max(s->gjobs[k].tlen, 0));
		}
		s->qbytes = qb; s->tbytes = tb;
		nq += qb; nt += tb;
	}
	q->gj.resize(total); q->gr.resize(total); q->gq.resize(nq ? nq : 1); q->gt.resize(nt ? nt : 1);
	size_t at = 0, qa = 0, ta = 0;
	for (Sub *s : grp) {
		if (s->qbytes) memcpy(q->gq.data() + qa, s->qpool, s->qbytes);
		if (s->tbytes) memcpy(q->gt.data() + ta, s->tpool, s->tbytes);
		for (int64_t k = 0; k < s->n; ++k) {
			ksw_b200_gjob_t j = s->gjobs[k];
			j.q_off += qa; j.t_off += ta;
			q->gj[at + (size_t)k] = j;
		}
		at += (size_t)s->n; qa += s->qbytes; ta += s->tbytes;
	}
	const uint32_t *pool = nullptr;
	int64_t n_cig = 0;
	const int rc = ksw_b200_global_batch(q->ctx, &grp[0]->cfg, (int64_t)total, q->gj.data(), q->gq.data(), q->gt.data(), q->gr.data(), &pool, &n_cig);
	at = 0;
	for (Sub *s : grp) {
		s->rc = rc;
		if (rc == 0) {
			int64_t ops = 0;
			for (int64_t k = 0; k < s->n; ++k) ops += q->gr[at + (size_t)k].n_cigar;
			s->cigar = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(ops > 0 ? ops : 1));
			int64_t w = 0;
			for (int64_t k = 0; k < s->n; ++k) {
				const ksw_b200_gres_t &r = q->gr[at + (size_t)k];
				s->gres[k].score = r.score; s->gres[k].n_cigar = r.n_cigar; s->gres[k].cigar_off = w;
				if (r.n_cigar) memcpy(s->cigar + w, pool + r.cigar_off, sizeof(uint32_t) * (size_t)r.n_cigar);
				w += r.n_cigar;
			}
			s->n_cigar = ops;
		}
		at += (size_t)s->n;
	}
}

void run_align(Server *q, std::vector<Sub *> &grp)
{
	size_t total = 0, nq = 0, nt = 0;
	for (Sub *s : grp) {
		total += (size_t)s->n;
		size_t qb = 0, tb = 0;
		for (int64_t k = 0; k < s->n; ++k) {
			qb = std::max<size_t>(qb, (size_t)s->ajobs[k].q_off + (size_t)std::max(s->ajobs[k].qlen, 0));
			tb = std::max<size_t>(tb, (size_t)s->ajobs[k].t_off + (size_t)std::max(s->ajobs[k].tlen, 0));
		}
		s->qbytes = qb; s->tbytes = tb;
		nq += qb; nt += tb;
	}
	q->aj.resize(total); q->ar.resize(total); q->gq.resize(nq ? nq : 1); q->gt.resize(nt ? nt : 1);
	size_t at = 0, qa = 0, ta = 0;
	for (Sub *s : grp) {
		if (s->qbytes) memcpy(q->gq.data() + qa, s->qpool, s->qbytes);
		if (s->tbytes) memcpy(q->gt.data() + ta, s->tpool, s->tbytes);
		for (int64_t k = 0; k < s->n; ++k) {
			ksw_b200_ajob_t j = s->ajobs[k];
			j.q_off += qa; j.t_off += ta;
			q->aj[at + (size_t)k] = j;
		}
		at += (size_t)s->n; qa += s->qbytes; ta += s->tbytes;
	}
	const int rc = ksw_b200_align_batch(q->ctx, &grp[0]->cfg, (int64_t)total, q->aj.data(), q->gq.data(), q->gt.data(), q->ar.data());
	at = 0;
	for (Sub *s : grp) {
		s->rc = rc;
		if (rc == 0) memcpy(s->ares, q->ar.data() + at, sizeof(ksw_b200_ares_t) * (size_t)s->n);
		at += (size_t)s->n;
	}
}

int submit(ksw_b200_queue *q, Sub &s)
{
	std::unique_lock<std::mutex> lk(q->mu);
	if (q->stop) return 1;
	q->pending.push_back(&s);
	std::vector<Sub *> grp;
	for (;;) {
		if (s.done) return s.rc;
		Server *lane = nullptr;
		if (!q->pending.empty()) for (Server &sv : q->srv) if (!sv.busy) { lane = &sv; break; }
		if (!lane) { s.cv.wait(lk); continue; }                       // served, or a lane freed up: look again
		// lead a batch: everything pending of the first submission's kind and scoring
		lane->busy = true;
		grp.clear();
		Sub *first = q->pending.front();
		for (auto it = q->pending.begin(); it != q->pending.end();) {
			if ((*it)->kind == first->kind && same_cfg((*it)->cfg, first->cfg)) { grp.push_back(*it); it = q->pending.erase(it); }
			else ++it;
		}
		if (!q->pending.empty())                                      // other work is left and the other lane may be free
			for (Server &sv : q->srv) if (!sv.busy) { q->pending.front()->cv.notify_one(); break; }
		lk.unlock();
		if (grp[0]->kind == 0) run_extend(lane, grp); else if (grp[0]->kind == 1) run_global(lane, grp); else run_align(lane, grp);
		lk.lock();
		if (grp[0]->rc) q->err = ksw_b200_strerror(lane->ctx);
		++q->n_batches; q->n_subs += (int64_t)grp.size();
		lane->busy = false;
		for (Sub *x : grp) { x->done = true; if (x != &s) x->cv.notify_one(); }
		if (!q->pending.empty()) q->pending.front()->cv.notify_one();   // the lane is free again: the oldest waiter may lead
		q->cv_idle.notify_all();
	}
}

} // namespace

extern "C" {

int ksw_b200_queue_create(int device, ksw_b200_queue_t **out)
{
	if (!out) return 1;
	*out = nullptr;
	ksw_b200_queue *q = new ksw_b200_queue();
	int lanes = KSW_QUEUE_LANES;
	if (const char *e = getenv("KSW_B200_QUEUE_LANES")) lanes = std::max(1, std::min(16, atoi(e)));
	q->srv.resize((size_t)lanes);
	for (Server &sv : q->srv) {
		const int rc = ksw_b200_ctx_create(device, &sv.ctx);
		if (rc) {
			for (Server &x : q->srv) if (x.ctx) ksw_b200_ctx_destroy(x.ctx);
			delete q;
			return rc;
		}
		ksw_b200_ctx_set_pack_threads(sv.ctx, 1);
	}
	*out = q;
	return 0;
}

void ksw_b200_queue_destroy(ksw_b200_queue_t *q)
{
	if (!q) return;
	{
		std::unique_lock<std::mutex> lk(q->mu);
		q->stop = true;                                              // no new submissions; the ones under way finish
		q->cv_idle.wait(lk, [&] {
			if (!q->pending.empty()) return false;
			for (Server &sv : q->srv) if (sv.busy) return false;
			return true;
		});
	}
	for (Server &sv : q->srv) ksw_b200_ctx_destroy(sv.ctx);
	delete q;
}

int ksw_b200_queue_ref_set(ksw_b200_queue_t *q, const uint8_t *pac, int64_t l_pac)
{
	if (!q) return 1;
	std::unique_lock<std::mutex> lk(q->mu);
	for (Server &sv : q->srv) {                                      // one copy per device: the second context shares it
		const int rc = ksw_b200_ref_set(sv.ctx, pac, l_pac);
		if (rc) { q->err = ksw_b200_strerror(sv.ctx); return rc; }
	}
	return 0;
}

int ksw_b200_queue_extend_ref(ksw_b200_queue_t *q, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_rjob_t *jobs,
                              const uint8_t *qpool, size_t qpool_bytes, ksw_b200_res_t *res)
{
	if (!q || !cfg || n < 0) return 1;
	if (n == 0) return 0;
	if (!jobs || !qpool || !res) return 1;
	Sub s;
	s.kind = 0; s.cfg = *cfg; s.n = n; s.rjobs = jobs; s.qpool = qpool; s.qbytes = qpool_bytes; s.res = res;
	return submit(q, s);
}

int ksw_b200_queue_global(ksw_b200_queue_t *q, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_gjob_t *jobs,
                          const uint8_t *qpool, const uint8_t *tpool, ksw_b200_gres_t *res,
                          uint32_t **cigar_pool, int64_t *n_cigar_total)
{
	if (!q || !cfg || n < 0 || !cigar_pool || !n_cigar_total) return 1;
	*cigar_pool = nullptr; *n_cigar_total = 0;
	if (n == 0) return 0;
	if (!jobs || !res) return 1;
	Sub s;
	s.kind = 1; s.cfg = *cfg; s.n = n; s.gjobs = jobs; s.qpool = qpool; s.tpool = tpool; s.gres = res;
	const int rc = submit(q, s);
	if (rc == 0) { *cigar_pool = s.cigar; *n_cigar_total = s.n_cigar; }
	else free(s.cigar);
	return rc;
}

int ksw_b200_queue_align(ksw_b200_queue_t *q, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_ajob_t *jobs,
                         const uint8_t *qpool, const uint8_t *tpool, ksw_b200_ares_t *res)
{
	if (!q || !cfg || n < 0) return 1;
	if (n == 0) return 0;
	if (!jobs || !qpool || !tpool || !res) return 1;
	Sub s;
	s.kind = 2; s.cfg = *cfg; s.n = n; s.ajobs = jobs; s.qpool = qpool; s.tpool = tpool; s.ares = res;
	return submit(q, s);
}

const char *ksw_b200_queue_strerror(const ksw_b200_queue_t *q) { return q ? q->err.c_str() : "null queue"; }

void ksw_b200_queue_stats(const ksw_b200_queue_t *q, int64_t *n_batches, int64_t *n_submissions)
{
	if (!q) return;
	if (n_batches) *n_batches = q->n_batches;
	if (n_submissions) *n_submissions = q->n_subs;
}

} // extern "C"
