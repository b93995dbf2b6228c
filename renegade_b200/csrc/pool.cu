// Prover pool: a FIFO job queue drained by N worker threads, each owning one context (stream +
// scratch) on the same device.  It is the device-side counterpart of the reference's
// `NativeProofManager` (crates/workers/proof-manager/src/implementations/native_proof_manager.rs:
// 138-201: a rayon pool fed with `spawn_fifo` from the job queue, one proof per worker): proving
// keys and SRS tables are shared read-only by all workers, so several proofs are in flight on one
// GPU and the latency-bound kernels of one proof (bucket reductions, scans, the host transcript)
// overlap the throughput-bound kernels of the others.
//
// Host-only code (std::thread) above the same C ABI the single-context entry points use.
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <thread>
#include <vector>

#include "b200prover.h"
#include "device_ctx.h"

namespace {

struct Bundle;

struct Job {
    enum Kind { kProve, kLink } kind = kProve;
    uint64_t ticket = 0;
    // prove
    const b200_pk* pk = nullptr;
    const uint64_t* wires = nullptr;
    std::vector<uint64_t> pub_inputs, blinders;
    b200_proof* proof = nullptr;
    uint64_t* link_poly = nullptr;
    // link
    const b200_bases* srs = nullptr;
    const uint64_t *a1 = nullptr, *a2 = nullptr;
    size_t len1 = 0, len2 = 0;
    uint64_t comm1[8], comm2[8];
    unsigned alignment = 0;
    size_t offset = 0, size = 0;
    b200_link_proof* link_proof = nullptr;
    std::shared_ptr<Bundle> bundle;  // set for the sub-jobs of a bundle (they carry no ticket of their own)
};

struct Done {
    int status = B200_OK;
    std::string message;
};

// A bundle: proofs first, then — once all of them are in — the link proofs between their wire-0 polynomials; one
// ticket for the whole thing.  The device-side shape of `NativeProofManager::handle_proof_job` for the settlement jobs
// (native_proof_manager.rs:526-584: prove, then `compute_*_link_proofs`, :726-782 forking the link proofs).
struct Bundle {
    uint64_t ticket = 0;
    const b200_bases* srs = nullptr;
    std::vector<b200_bundle_proof> proofs;
    std::vector<b200_bundle_link> links;
    size_t pending = 0;     // sub-jobs of the current phase still queued or running
    bool linking = false;   // false: proof phase, true: link phase
    Done result;            // first failure wins
};

}  // namespace

struct b200_pool {
    std::vector<b200_ctx*> ctxs;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<std::unique_ptr<Job>> queue;
    std::map<uint64_t, Done> done;   // finished, not yet waited for
    std::set<uint64_t> inflight;     // queued or running
    uint64_t next_ticket = 1;
    uint64_t n_submitted = 0, n_completed = 0, n_failed = 0, n_running = 0;
    bool stopping = false;

    static std::unique_ptr<Job> make_link_job(const std::shared_ptr<Bundle>& b, const b200_bundle_link& l) {
        const b200_bundle_proof& pa = b->proofs[l.a];
        const b200_bundle_proof& pb = b->proofs[l.b];
        std::unique_ptr<Job> j(new Job());
        j->kind = Job::kLink;
        j->bundle = b;
        j->srs = b->srs;
        j->a1 = pa.link_poly;
        j->len1 = ((size_t)1 << b200_pk_log_n(pa.pk)) + 2;
        j->a2 = pb.link_poly;
        j->len2 = ((size_t)1 << b200_pk_log_n(pb.pk)) + 2;
        std::memcpy(j->comm1, pa.proof->wires_poly_comms[0], sizeof j->comm1);  // `linking_wire_comm` of each hint
        std::memcpy(j->comm2, pb.proof->wires_poly_comms[0], sizeof j->comm2);
        j->alignment = l.alignment;
        j->offset = l.offset;
        j->size = l.size;
        j->link_proof = l.proof;
        return j;
    }

    void run(unsigned w) {
        for (;;) {
            std::unique_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stopping || !queue.empty(); });
                if (queue.empty()) return;  // stopping and drained
                job = std::move(queue.front());
                queue.pop_front();
                ++n_running;
            }
            Done d;
            if (job->kind == Job::kProve) {
                d.status = b200_plonk_prove(ctxs[w], job->pk, job->wires,
                                            job->pub_inputs.empty() ? nullptr : job->pub_inputs.data(),
                                            job->blinders.data(), job->proof, job->link_poly, nullptr);
            } else {
                d.status = b200_plonk_link(ctxs[w], job->srs, job->a1, job->len1, job->a2, job->len2,
                                           job->comm1, job->comm2, job->alignment, job->offset, job->size,
                                           job->link_proof, nullptr);
            }
            if (d.status != B200_OK) d.message = b200_last_error();  // this worker's thread-local message
            bool new_jobs = false;
            {
                std::lock_guard<std::mutex> lk(mu);
                --n_running;
                ++n_completed;
                if (d.status != B200_OK) ++n_failed;
                if (!job->bundle) {
                    inflight.erase(job->ticket);
                    done.emplace(job->ticket, std::move(d));
                } else {
                    Bundle& b = *job->bundle;
                    if (d.status != B200_OK && b.result.status == B200_OK) b.result = d;
                    if (--b.pending == 0) {
                        if (!b.linking && b.result.status == B200_OK && !b.links.empty()) {
                            b.linking = true;  // every proof is in: fork the link proofs
                            for (const b200_bundle_link& l : b.links) queue.push_back(make_link_job(job->bundle, l));
                            b.pending = b.links.size();
                            n_submitted += b.links.size();
                            new_jobs = true;
                        } else {
                            inflight.erase(b.ticket);
                            done.emplace(b.ticket, b.result);
                        }
                    }
                }
            }
            if (new_jobs) cv_job.notify_all();
            cv_done.notify_all();
        }
    }
};

extern "C" {

int b200_pool_create(int device, unsigned n_workers, b200_pool** out) {
    B200_TRY
    if (!out || n_workers == 0 || n_workers > 64) {
        b200::set_error("pool: n_workers must be in [1, 64]");
        return B200_ERR_INVALID;
    }
    std::unique_ptr<b200_pool> p(new b200_pool());
    for (unsigned i = 0; i < n_workers; ++i) {
        b200_ctx* c = nullptr;
        const int rc = b200_init(device, &c);
        if (rc != B200_OK) {
            for (b200_ctx* q : p->ctxs) b200_shutdown(q);
            return rc;
        }
        p->ctxs.push_back(c);
    }
    b200_pool* raw = p.release();
    for (unsigned i = 0; i < n_workers; ++i) raw->workers.emplace_back([raw, i] { raw->run(i); });
    *out = raw;
    return B200_OK;
    B200_CATCH
}

void b200_pool_destroy(b200_pool* pool) {
    if (!pool) return;
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        pool->stopping = true;  // workers drain what is queued, then leave
    }
    pool->cv_job.notify_all();
    for (std::thread& t : pool->workers) t.join();
    for (b200_ctx* c : pool->ctxs) b200_shutdown(c);
    delete pool;
}

unsigned b200_pool_workers(const b200_pool* pool) { return pool ? (unsigned)pool->ctxs.size() : 0u; }

b200_ctx* b200_pool_ctx(b200_pool* pool, unsigned worker) {
    if (!pool || worker >= pool->ctxs.size()) return nullptr;
    return pool->ctxs[worker];
}

static int pool_push(b200_pool* pool, std::unique_ptr<Job> job, uint64_t* ticket) {
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        if (pool->stopping) {
            b200::set_error("pool: shutting down");
            return B200_ERR_INVALID;
        }
        job->ticket = pool->next_ticket++;
        *ticket = job->ticket;
        ++pool->n_submitted;
        pool->inflight.insert(job->ticket);
        pool->queue.push_back(std::move(job));
    }
    pool->cv_job.notify_one();
    return B200_OK;
}

int b200_pool_submit_prove(b200_pool* pool, const b200_pk* pk, const uint64_t* wires,
                           const uint64_t* pub_inputs, size_t num_inputs, const uint64_t* blinders,
                           b200_proof* proof, uint64_t* link_poly, uint64_t* ticket) {
    B200_TRY
    if (!pool || !pk || !wires || !blinders || !proof || !ticket || (num_inputs && !pub_inputs)) {
        b200::set_error("pool_submit_prove: null argument");
        return B200_ERR_INVALID;
    }
    if (num_inputs != b200_pk_num_inputs(pk)) {  // the worker reads exactly pk.num_inputs elements
        b200::set_error("pool_submit_prove: num_inputs does not match the proving key");
        return B200_ERR_INVALID;
    }
    std::unique_ptr<Job> j(new Job());
    j->kind = Job::kProve;
    j->pk = pk;
    j->wires = wires;
    j->pub_inputs.assign(pub_inputs, pub_inputs + 4 * num_inputs);
    j->blinders.assign(blinders, blinders + 4 * 17);
    j->proof = proof;
    j->link_poly = link_poly;
    return pool_push(pool, std::move(j), ticket);
    B200_CATCH
}

int b200_pool_submit_link(b200_pool* pool, const b200_bases* srs, const uint64_t* a1, size_t len1,
                          const uint64_t* a2, size_t len2, const uint64_t* comm1, const uint64_t* comm2,
                          unsigned alignment, size_t offset, size_t size, b200_link_proof* proof,
                          uint64_t* ticket) {
    B200_TRY
    if (!pool || !srs || !a1 || !a2 || !comm1 || !comm2 || !proof || !ticket) {
        b200::set_error("pool_submit_link: null argument");
        return B200_ERR_INVALID;
    }
    std::unique_ptr<Job> j(new Job());
    j->kind = Job::kLink;
    j->srs = srs;
    j->a1 = a1;
    j->len1 = len1;
    j->a2 = a2;
    j->len2 = len2;
    std::memcpy(j->comm1, comm1, sizeof j->comm1);
    std::memcpy(j->comm2, comm2, sizeof j->comm2);
    j->alignment = alignment;
    j->offset = offset;
    j->size = size;
    j->link_proof = proof;
    return pool_push(pool, std::move(j), ticket);
    B200_CATCH
}

int b200_pool_submit_bundle(b200_pool* pool, const b200_bases* srs, const b200_bundle_proof* proofs, size_t n_proofs,
                            const b200_bundle_link* links, size_t n_links, uint64_t* ticket) {
    B200_TRY
    if (!pool || !proofs || n_proofs == 0 || (n_links && (!links || !srs)) || !ticket) {
        b200::set_error("pool_submit_bundle: null argument");
        return B200_ERR_INVALID;
    }
    for (size_t i = 0; i < n_proofs; ++i) {
        const b200_bundle_proof& p = proofs[i];
        if (!p.pk || !p.wires || !p.blinders || !p.proof || (p.num_inputs && !p.pub_inputs) ||
            p.num_inputs != b200_pk_num_inputs(p.pk)) {
            b200::set_error("pool_submit_bundle: bad proof entry (null pointer or num_inputs does not match the key)");
            return B200_ERR_INVALID;
        }
    }
    for (size_t i = 0; i < n_links; ++i) {
        const b200_bundle_link& l = links[i];
        if (l.a >= n_proofs || l.b >= n_proofs || !l.proof || !proofs[l.a].link_poly || !proofs[l.b].link_poly) {
            b200::set_error("pool_submit_bundle: a link names a proof that is missing or has no link_poly buffer");
            return B200_ERR_INVALID;
        }
    }
    std::shared_ptr<Bundle> b(new Bundle());
    b->srs = srs;
    b->proofs.assign(proofs, proofs + n_proofs);
    b->links.assign(links, links + n_links);
    b->pending = n_proofs;
    std::vector<std::unique_ptr<Job>> jobs;
    for (size_t i = 0; i < n_proofs; ++i) {
        const b200_bundle_proof& p = proofs[i];
        std::unique_ptr<Job> j(new Job());
        j->kind = Job::kProve;
        j->bundle = b;
        j->pk = p.pk;
        j->wires = p.wires;
        j->pub_inputs.assign(p.pub_inputs, p.pub_inputs + 4 * p.num_inputs);
        j->blinders.assign(p.blinders, p.blinders + 4 * 17);
        j->proof = p.proof;
        j->link_poly = p.link_poly;
        jobs.push_back(std::move(j));
    }
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        if (pool->stopping) {
            b200::set_error("pool: shutting down");
            return B200_ERR_INVALID;
        }
        b->ticket = pool->next_ticket++;
        *ticket = b->ticket;
        pool->n_submitted += n_proofs;
        pool->inflight.insert(b->ticket);
        for (auto& j : jobs) pool->queue.push_back(std::move(j));
    }
    pool->cv_job.notify_all();
    return B200_OK;
    B200_CATCH
}

int b200_pool_wait(b200_pool* pool, uint64_t ticket) {
    B200_TRY
    if (!pool || ticket == 0) {
        b200::set_error("pool_wait: bad ticket");
        return B200_ERR_INVALID;
    }
    std::unique_lock<std::mutex> lk(pool->mu);
    if (ticket >= pool->next_ticket) {
        b200::set_error("pool_wait: unknown ticket");
        return B200_ERR_INVALID;
    }
    for (;;) {
        auto it = pool->done.find(ticket);
        if (it != pool->done.end()) {
            const Done d = std::move(it->second);
            pool->done.erase(it);
            lk.unlock();
            if (d.status != B200_OK) b200::set_error(d.message);
            return d.status;
        }
        if (!pool->inflight.count(ticket)) {
            b200::set_error("pool_wait: ticket already waited for");
            return B200_ERR_INVALID;
        }
        pool->cv_done.wait(lk);
    }
    B200_CATCH
}

int b200_pool_wait_all(b200_pool* pool) {
    B200_TRY
    if (!pool) {
        b200::set_error("pool_wait_all: null pool");
        return B200_ERR_INVALID;
    }
    std::unique_lock<std::mutex> lk(pool->mu);
    pool->cv_done.wait(lk, [&] { return pool->queue.empty() && pool->n_running == 0; });
    int status = B200_OK;
    std::string message;
    for (auto& kv : pool->done)  // tickets are increasing: report the oldest failure
        if (kv.second.status != B200_OK && status == B200_OK) {
            status = kv.second.status;
            message = kv.second.message;
        }
    pool->done.clear();
    lk.unlock();
    if (status != B200_OK) b200::set_error(message);
    return status;
    B200_CATCH
}

int b200_pool_stats(b200_pool* pool, uint64_t out[4]) {
    if (!pool || !out) {
        b200::set_error("pool_stats: null argument");
        return B200_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lk(pool->mu);
    out[0] = pool->n_submitted;
    out[1] = pool->n_completed;
    out[2] = pool->n_failed;
    out[3] = pool->queue.size();
    return B200_OK;
}

}  // extern "C"
