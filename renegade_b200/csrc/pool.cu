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

// ---------------------------------------------------------------------------------------------------------------
// The whole box behind one object: one prover pool per device, the SRS tables and proving keys replicated on every
// device (whole proofs do not shard: SURVEY.md §8(e) "independent jobs => replicas, one proof stream per GPU, no
// communication"), each job routed to the device with the fewest unfinished jobs.  What a single Rust host process
// (the relayer's `NativeProofManager`) needs to use all GPUs without managing devices itself.
// ---------------------------------------------------------------------------------------------------------------
struct b200_box {
    std::vector<int> devices;
    std::vector<b200_pool*> pools;
};
struct b200_box_srs {
    std::vector<b200_bases*> per_dev;
};
struct b200_box_pk {
    std::vector<b200_pk*> per_dev;
};

namespace {
constexpr int kBoxTicketShift = 48;  // ticket = device index << 48 | the device pool's ticket

int box_least_loaded(b200_box* box) {
    int best = 0;
    uint64_t best_load = ~0ull;
    for (size_t i = 0; i < box->pools.size(); ++i) {
        uint64_t st[4];
        if (b200_pool_stats(box->pools[i], st) != B200_OK) continue;
        const uint64_t load = st[0] - st[1];  // submitted - completed
        if (load < best_load) {
            best_load = load;
            best = (int)i;
        }
    }
    return best;
}
}  // namespace

extern "C" {

void b200_box_destroy(b200_box* box) {
    if (!box) return;
    for (b200_pool* p : box->pools) b200_pool_destroy(p);
    delete box;
}

int b200_box_create(const int* devices, int n_dev, unsigned workers_per_device, b200_box** out) {
    B200_TRY
    if (!devices || n_dev <= 0 || n_dev > 64 || !out) {
        b200::set_error("box_create: need 1..64 device ordinals");
        return B200_ERR_INVALID;
    }
    std::unique_ptr<b200_box> box(new b200_box());
    for (int i = 0; i < n_dev; ++i) {
        b200_pool* p = nullptr;
        const int rc = b200_pool_create(devices[i], workers_per_device, &p);
        if (rc != B200_OK) {
            b200_box_destroy(box.release());
            return rc;
        }
        box->devices.push_back(devices[i]);
        box->pools.push_back(p);
    }
    *out = box.release();
    return B200_OK;
    B200_CATCH
}

int b200_box_devices(const b200_box* box) { return box ? (int)box->pools.size() : 0; }
b200_pool* b200_box_pool(b200_box* box, int i) {
    if (!box || i < 0 || i >= (int)box->pools.size()) return nullptr;
    return box->pools[i];
}

void b200_box_srs_free(b200_box* box, b200_box_srs* srs) {
    if (!srs) return;
    for (size_t i = 0; i < srs->per_dev.size(); ++i)
        if (srs->per_dev[i]) b200_bases_free(box && i < box->pools.size() ? b200_pool_ctx(box->pools[i], 0) : nullptr, srs->per_dev[i]);
    delete srs;
}

int b200_box_srs_load(b200_box* box, const uint8_t* points64, size_t n, int window_bits, int check_on_curve, b200_box_srs** out) {
    B200_TRY
    if (!box || !points64 || !out) return B200_ERR_INVALID;
    std::unique_ptr<b200_box_srs> s(new b200_box_srs());
    s->per_dev.assign(box->pools.size(), nullptr);
    for (size_t i = 0; i < box->pools.size(); ++i) {
        // the on-curve assertion of srs.rs:178-179 once is enough: every device gets the same bytes
        const int rc = b200_bases_load(b200_pool_ctx(box->pools[i], 0), points64, n, window_bits, i == 0 ? check_on_curve : 0, &s->per_dev[i]);
        if (rc != B200_OK) {
            b200_box_srs_free(box, s.release());
            return rc;
        }
    }
    *out = s.release();
    return B200_OK;
    B200_CATCH
}

void b200_box_pk_free(b200_box* box, b200_box_pk* pk) {
    if (!pk) return;
    for (size_t i = 0; i < pk->per_dev.size(); ++i)
        if (pk->per_dev[i]) b200_pk_free(box && i < box->pools.size() ? b200_pool_ctx(box->pools[i], 0) : nullptr, pk->per_dev[i]);
    delete pk;
}

int b200_box_preprocess(b200_box* box, const b200_box_srs* srs, unsigned log_n, size_t num_inputs, const uint64_t* selectors_evals,
                        const uint64_t* perm, const uint64_t* k, b200_box_pk** out) {
    B200_TRY
    if (!box || !srs || !out || srs->per_dev.size() != box->pools.size()) return B200_ERR_INVALID;
    std::unique_ptr<b200_box_pk> pk(new b200_box_pk());
    pk->per_dev.assign(box->pools.size(), nullptr);
    for (size_t i = 0; i < box->pools.size(); ++i) {
        const int rc = b200_plonk_preprocess(b200_pool_ctx(box->pools[i], 0), srs->per_dev[i], log_n, num_inputs, selectors_evals, perm, k,
                                             &pk->per_dev[i]);
        if (rc != B200_OK) {
            b200_box_pk_free(box, pk.release());
            return rc;
        }
    }
    *out = pk.release();
    return B200_OK;
    B200_CATCH
}

int b200_box_pk_verifying_key(const b200_box_pk* pk, uint64_t* selector_comms, uint64_t* sigma_comms) {
    if (!pk || pk->per_dev.empty()) return B200_ERR_INVALID;
    return b200_pk_verifying_key(pk->per_dev[0], selector_comms, sigma_comms);  // identical on every device
}

int b200_box_submit_prove(b200_box* box, const b200_box_pk* pk, const uint64_t* wires, const uint64_t* pub_inputs, size_t num_inputs,
                          const uint64_t* blinders, b200_proof* proof, uint64_t* link_poly, uint64_t* ticket) {
    B200_TRY
    if (!box || !pk || !ticket || pk->per_dev.size() != box->pools.size()) {
        b200::set_error("box_submit_prove: null argument or key of another box");
        return B200_ERR_INVALID;
    }
    const int d = box_least_loaded(box);
    uint64_t t = 0;
    const int rc = b200_pool_submit_prove(box->pools[d], pk->per_dev[d], wires, pub_inputs, num_inputs, blinders, proof, link_poly, &t);
    if (rc != B200_OK) return rc;
    *ticket = ((uint64_t)d << kBoxTicketShift) | t;
    return B200_OK;
    B200_CATCH
}

int b200_box_submit_bundle(b200_box* box, const b200_box_srs* srs, const b200_box_pk* const* pks, const b200_bundle_proof* proofs,
                           size_t n_proofs, const b200_bundle_link* links, size_t n_links, uint64_t* ticket) {
    B200_TRY
    if (!box || !pks || !proofs || n_proofs == 0 || !ticket || (n_links && !srs)) {
        b200::set_error("box_submit_bundle: null argument");
        return B200_ERR_INVALID;
    }
    const int d = box_least_loaded(box);  // a bundle stays on one device: its link proofs need its proofs' polynomials
    std::vector<b200_bundle_proof> local(proofs, proofs + n_proofs);
    for (size_t i = 0; i < n_proofs; ++i) {
        if (!pks[i] || pks[i]->per_dev.size() != box->pools.size()) {
            b200::set_error("box_submit_bundle: missing key");
            return B200_ERR_INVALID;
        }
        local[i].pk = pks[i]->per_dev[d];
    }
    uint64_t t = 0;
    const int rc = b200_pool_submit_bundle(box->pools[d], srs ? srs->per_dev[d] : nullptr, local.data(), n_proofs, links, n_links, &t);
    if (rc != B200_OK) return rc;
    *ticket = ((uint64_t)d << kBoxTicketShift) | t;
    return B200_OK;
    B200_CATCH
}

int b200_box_wait(b200_box* box, uint64_t ticket) {
    B200_TRY
    const uint64_t d = ticket >> kBoxTicketShift;
    if (!box || d >= box->pools.size()) {
        b200::set_error("box_wait: bad ticket");
        return B200_ERR_INVALID;
    }
    return b200_pool_wait(box->pools[d], ticket & (((uint64_t)1 << kBoxTicketShift) - 1));
    B200_CATCH
}

int b200_box_ticket_device(const b200_box* box, uint64_t ticket) {
    const uint64_t d = ticket >> kBoxTicketShift;
    return box && d < box->pools.size() ? (int)d : -1;
}

}  // extern "C"
