/*
 * scalesim_b200.h — C ABI of libscalesim_b200.so: the B200-native Time Warp core.
 *
 * This is the drop-in boundary for ONE path of masatoshihanai/ScaleSim: the optimistic PDES
 * core that scalesim::runner<App> drives (reference include/scalesim/simulation/runner.hpp).
 * The reference has no FFI; its boundary is the C++ template surface. The host-side mirror of
 * that surface (include/scalesim/*.hpp in this repo) talks to the device engine only through
 * the functions below. Plain C: int status returns (0 = ok), no exceptions cross the boundary,
 * the engine owns all device memory, the caller owns every host buffer it passes, one host
 * thread per engine. There is NO CPU fallback: ssb_create fails when no CUDA device is usable.
 *
 * Which reference interface each entry point replaces (file:line in the reference tree):
 *   ssb_create / ssb_destroy      runner<App>::runner, lp_mngr::init_lps / delete_lps
 *                                 (runner.hpp:31-35, logical_process.hpp:265-287, 305-313)
 *   ssb_set_partition             lp_mngr::init_partition(App::init_partition_index(ranks))
 *                                 (runner.hpp:98, logical_process.hpp:258-263)
 *   ssb_set_model_data            App::init() — PHOLD's file-static tables (phold.hpp:144-164),
 *                                 TrafficSim's routes carried inside events (traffic_sim.hpp:43)
 *   ssb_load_states               App::init_states_in_this_rank -> lp::init_state (runner.hpp:108-118)
 *   ssb_load_events               App::init_events -> shuffle -> lp::init_event = lp::buffer
 *                                 (runner.hpp:131-147, logical_process.hpp:115-127)
 *   ssb_run / ssb_step            runner::loop — run_lp/run_lp_aggr, GVT, std_out, clear
 *                                 (runner.hpp:350-396, 517-583)
 *   ssb_committed_digest,         eventq::std_out + sim_event::result_out — the "RE,..." lines
 *   ssb_drain_committed*,         (queue.hpp:203-211, sim_obj.hpp:66-77); the reference prints them while it simulates
 *   ssb_run_stream_packed         (runner.hpp:369-380)
 *   ssb_drain_history,            lp::clear_old_ev / clear_old_st -> eventq::store_* / stateq::store_state ->
 *   ssb_run_stream_history        leveldb_store::put (logical_process.hpp:187-203, queue.hpp:179-201, 304-323,
 *                                 leveldb_store.hpp:132-185)
 *   ssb_load_history              lp::flush_buf lazy load -> load_events / load_cancels / load_prev_st
 *                                 (logical_process.hpp:132-153, queue.hpp:213-241, 325-331)
 *   ssb_get_stats, ssb_kernel_times  runner::finish report + the stopwatch phases (runner.hpp:482-507,
 *                                 util/stopwatch.hpp)
 *   ssb_reset                     (no reference equivalent: the reference is one simulation per process)
 *   ssb_read_states               (no reference equivalent: sim_state::result_out is empty,
 *                                 sim_obj.hpp:97; used by the parity tests for final LP states)
 *   ssb_comm_*                    event_com / mpi_thr / mpi_gsync (com/*.hpp) — replaced by peer-memory mailboxes over
 *                                 NVLink (events + GVT min-reduction), or NCCL send/recv + allreduce-min, between
 *                                 one-engine-per-GPU processes
 */
#ifndef SCALESIM_B200_H_
#define SCALESIM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_ABI_VERSION 2

typedef struct ssb_engine ssb_engine;

/* device twins of the application handlers (App::event_handler) */
enum ssb_model_id {
  SSB_MODEL_PHOLD = 1,    /* src/phold/phold.hpp:256-281 */
  SSB_MODEL_TRAFFIC = 2,  /* src/trafficsim/traffic_sim.hpp:279-342 */
  SSB_MODEL_PHOLD_STATEFUL = 3 /* test model: PHOLD whose LP state mutates and feeds back into events */
};

/* status codes */
enum ssb_status {
  SSB_OK = 0,
  SSB_ERR_INVALID = 1,      /* bad argument / call order */
  SSB_ERR_CUDA = 2,         /* CUDA runtime failure (incl. no device) */
  SSB_ERR_RANGE = 3,        /* a value does not fit the device record (see limits below) */
  SSB_ERR_CAPACITY = 4,     /* a device pool overflowed (max_events / max_batch / ...) */
  SSB_ERR_MODEL = 5,        /* model data missing or inconsistent */
  SSB_ERR_COMM = 6          /* NCCL failure, or a peer partition did not reach the GVT exchange */
};

/*
 * Device record limits (asserted at load and at every emit; violation -> SSB_ERR_RANGE):
 *   0 <= time < 2^31, 0 <= event id < 2^32-1, 0 <= LP id < 2^32-1, source may be -1.
 * The sort key is (receive_time, id) packed into one u64 — ScaleSim's timestamp order
 * (util/timestamp.hpp:53-59).
 */
typedef struct ssb_config {
  int32_t abi_version;      /* SSB_ABI_VERSION */
  int32_t model;            /* ssb_model_id */
  int32_t device;           /* CUDA device ordinal */
  int32_t rank, world;      /* this partition / number of partitions (one engine per GPU) */
  int32_t window_buckets;   /* optimistic window, in calendar buckets (>= 1) */
  int64_t n_lp;             /* global number of LPs; LP ids are [0, n_lp) */
  int64_t finish_time;      /* App::finish_time() */
  int64_t bucket_ticks;     /* calendar bucket width in simulation ticks (>= 1) */
  int64_t max_events;       /* capacity of the device event pool (live + uncommitted events) */
  int64_t max_batch;        /* max events integrated per round */
  int64_t max_committed;    /* capacity of the device committed-record log; 0 = digest only */
  int64_t max_snapshots;    /* capacity of the state-snapshot ring */
  int32_t rounds_per_sync;  /* rounds enqueued between host checks of the done flag */
  int32_t flags;            /* SSB_FLAG_* */
} ssb_config;

#define SSB_FLAG_DEBUG_CHECKS 1   /* extra device-side invariant checks */
#define SSB_FLAG_KERNEL_TIMING 2  /* bracket every kernel launch with CUDA events (see ssb_kernel_times) */
#define SSB_FLAG_HISTORY 4        /* --diff_init: keep the sent-log entry of every committed event for the history store
                                     (set at ssb_create; needs max_committed > 0) */

/* Host-side POD mirror of the App::Event accessors (sim_obj.hpp:50-59) + model payload. */
typedef struct ssb_event {
  int64_t id;            /* Event::id() */
  int64_t src;           /* Event::source() (-1 = none) */
  int64_t dst;           /* Event::destination() */
  int64_t recv_time;     /* Event::receive_time() */
  int64_t send_time;     /* Event::send_time() */
  int64_t p0;            /* PHOLD: num_hops(); TrafficSim: offset of destination() in the route pool */
  int64_t p1;            /* TrafficSim: number of route nodes left, including destination() */
  int64_t flags;         /* bit 0: is_cancel() */
} ssb_event;

/* One committed result = one "RE,id,src,send,dst,recv[,START|,END]" line (sim_obj.hpp:66-77). */
typedef struct ssb_record {
  int64_t id, src, send_time, dst, recv_time;
  int64_t tag;           /* 0 none, 1 START, 2 END */
} ssb_record;

/* Compact committed record as it sits in the device log (32 bytes; what ssb_drain_committed_raw copies).
 * recv_time = key >> 32, id = key & 0xffffffff; src 0xffffffff = -1; tag as in ssb_record. */
typedef struct ssb_raw_record {
  uint64_t key;
  uint32_t dst, src, send_time, p0, p1, tag;
} ssb_raw_record;

/* Sent-log entry of one committed event (what eventq::set_cancel keeps, queue.hpp:151-157): the event it produced.
 * dst 0xffffffff = the handler produced nothing. Travels next to the ssb_raw_record of the event in the
 * exact-differential history store (the reference's event / cancel / state LevelDBs, leveldb_store.hpp). */
typedef struct ssb_sent_record {
  uint32_t dst;          /* destination LP of the produced event */
  uint32_t id;           /* its id */
  uint32_t recv_time;    /* its receive_time */
  uint32_t reserved;
} ssb_sent_record;

/* Packed committed record (24 bytes; what ssb_drain_committed_packed copies): the six fields of one "RE," line in
 * 32 bits each. src 0xffffffff = -1. */
typedef struct ssb_packed_record {
  uint32_t id, src, send_time, dst, recv_time, tag;
} ssb_packed_record;

typedef struct ssb_stats {
  int64_t rounds;            /* windows executed */
  int64_t processed;         /* handler invocations ("# of processing events", runner.hpp:495) */
  int64_t rollbacks;         /* LP turns that rolled back */
  int64_t undone;            /* processed events undone by rollbacks */
  int64_t antis_sent;        /* anti-messages emitted ("# of cancels", runner.hpp:496) */
  int64_t annihilated;       /* positive/anti pairs annihilated */
  int64_t duplicates;        /* duplicate keys ignored (queue.hpp:92) */
  int64_t committed;         /* "# of output events" (runner.hpp:497) */
  int64_t beyond_finish;     /* events emitted with receive_time >= finish_time (never executed) */
  int64_t remote_sent;       /* events sent to other partitions */
  int64_t gvt_time, gvt_id;  /* final GVT */
  int64_t pool_high_water;   /* peak event-pool slots in use */
  int64_t log_dropped;       /* committed records not kept because max_committed was reached */
  double loop_ms;            /* device time of the window loop (CUDA events) */
  int64_t fix_rounds;        /* rounds in which some LP's turn was not plain (the k_fix_* kernels ran) */
  int64_t fix_turns;         /* irregular LP turns (anti-message, duplicate key, straggler, state-mutating handler) */
  int64_t exchanges;         /* global synchronisations (GVT + mailbox exchange); = rounds unless ssb_set_sync_interval > 1 */
} ssb_stats;

int ssb_abi_version(void);
const char* ssb_last_error(ssb_engine* e);      /* e may be NULL: error of the last failed ssb_create */

int ssb_create(const ssb_config* cfg, ssb_engine** out);
void ssb_destroy(ssb_engine* e);

/* lp_to_part[lp] = owning partition in [0, world). NULL = PHOLD's default `lp % world` (phold.hpp:181-186). */
int ssb_set_partition(ssb_engine* e, const int64_t* lp_to_part, int64_t n_lp);

/*
 * Model data (host pointers, copied):
 *   PHOLD   slot 0: LATENCY_TABLE  int64[n]   slot 1: REMOTE_COM_TABLE int32[n]   (phold.hpp:56-58)
 *           slot 2: int64[2] = { LOOK_AHEAD, NUM_LP }
 *   TRAFFIC slot 0: route pool int64[n] (all route nodes, concatenated; events index into it)
 *           slot 1: road table int64[3k] = {destination, speed km/h, length m} per road, for the junctions with more
 *                   than 2 out-roads (state word 2 = index of the junction's first road; see the state layout below)
 * TRAFFIC LP state (8 words): [0] id, [1] n = number of out-roads; n <= 2: words 2.. = n x {dst, speed, length} inline;
 *   n > 2: word 2 = index (in roads) of the LP's first road in the slot-1 table, its n roads contiguous there.
 */
int ssb_set_model_data(ssb_engine* e, int slot, const void* data, size_t bytes);

/* states: n records of `ssb_state_words()` 32-bit words each, for LPs lp_ids[i] owned by this rank. */
int ssb_state_words(ssb_engine* e);
int ssb_load_states(ssb_engine* e, const int64_t* lp_ids, const uint32_t* words, int64_t n);
int ssb_read_states(ssb_engine* e, const int64_t* lp_ids, uint32_t* words, int64_t n);

/* initial (or injected) events; events whose destination is owned by another rank are forwarded at the next step. */
int ssb_load_events(ssb_engine* e, const ssb_event* ev, int64_t n);

/* Same as ssb_load_events for events already in the compact 32-byte layout (half the host-to-device bytes):
 * key = receive_time << 32 | id, src 0xffffffff = -1, tag bit 0 = is_cancel(). `ev` may be pinned. */
int ssb_load_events_raw(ssb_engine* e, const ssb_raw_record* ev, int64_t n);

/* Run windows until GVT.time >= finish_time (runner.hpp:392). */
int ssb_run(ssb_engine* e, ssb_stats* stats);
/* Run exactly one window (one round). *done = 1 when GVT.time >= finish_time. */
int ssb_step(ssb_engine* e, int* done);

int ssb_get_stats(ssb_engine* e, ssb_stats* stats);
/* digest of the committed multiset: d[0]=count, d[1]=sum h, d[2]=xor h, d[3]=sum mix(h+K) */
int ssb_committed_digest(ssb_engine* e, uint64_t d[4]);
/* copy up to cap committed records to host memory and remove them from the device log */
int ssb_drain_committed(ssb_engine* e, ssb_record* out, int64_t cap, int64_t* n);

/* same, without widening: n raw 32-byte records straight from the device log (one D2H copy; `out` may be pinned) */
int ssb_drain_committed_raw(ssb_engine* e, ssb_raw_record* out, int64_t cap, int64_t* n);

/* same, 24-byte records (packed on the device first: three quarters of the device-to-host bytes of the raw form) */
int ssb_drain_committed_packed(ssb_engine* e, ssb_packed_record* out, int64_t cap, int64_t* n);

/* ssb_run + ssb_drain_committed_packed in one call, overlapped: while the window loop runs on the device, the
 * committed records of the rounds that are already complete are packed and copied to `out` on a second stream
 * (the reference prints its result lines while it simulates, runner.hpp:369-380). Returns when the simulation has
 * finished and every record is in `out` (*n of them; cap must hold them all). */
int ssb_run_drain_packed(ssb_engine* e, ssb_packed_record* out, int64_t cap, int64_t* n, ssb_stats* stats);

/* Streaming forms of the two calls above and of ssb_drain_history: the records leave through two bounded pinned
 * staging buffers owned by the engine (2 x 2M records) while the window loop runs, and are handed to `sink` chunk
 * by chunk on the calling thread (the next chunk is already crossing PCIe meanwhile) — host memory stays bounded
 * whatever the run commits. This is what the reference's store does NOT do (leveldb_store.hpp:132-141 keeps every
 * put in one in-memory WriteBatch until finish()). A sink returns 0 to go on; anything else
 * stops the draining and the call fails with SSB_ERR_INVALID once the loop has ended. The device log must hold the
 * whole run (max_committed >= committed events): it is not a ring, the device is never throttled by the host.
 * ssb_run_stream_history needs SSB_FLAG_HISTORY. */
typedef int (*ssb_packed_sink)(void* user, const ssb_packed_record* rec, int64_t n);
typedef int (*ssb_history_sink)(void* user, const ssb_raw_record* rec, const ssb_sent_record* sent, int64_t n);
int ssb_run_stream_packed(ssb_engine* e, ssb_packed_sink sink, void* user, ssb_stats* stats);
int ssb_run_stream_history(ssb_engine* e, ssb_history_sink sink, void* user, ssb_stats* stats);

/* ---- exact-differential simulation (reference --diff_init / --diff_repeat; runner.hpp:178-348, queue.hpp:179-241,
 * logical_process.hpp:132-153, leveldb_store.hpp) ----
 * --diff_init: with SSB_FLAG_HISTORY, ssb_drain_history hands out every committed event together with its sent-log
 * entry — the logical content of the reference's three stores (event: key (dst, recv_time, id); cancel and state:
 * key (dst, recv_time, id) of the events that produced something, plus one (-1,-1) state per LP). `rec` / `sent` may
 * be pinned host buffers; the copies are asynchronous device-to-host transfers.
 * --diff_repeat: ssb_load_history puts a recorded history back as PROCESSED events (calendar, processed chains,
 * sent-log) without executing anything; what-if events are then injected with ssb_load_events and ssb_run re-executes
 * exactly what they invalidate: a touched LP rolls back to the query time, its recorded outputs go out as
 * anti-messages and cascade — the reference's lazy history load (R13), with the history already resident in HBM.
 * Only events (re-)executed by the repeat run are committed (output / digest). */
int ssb_drain_history(ssb_engine* e, ssb_raw_record* rec, ssb_sent_record* sent, int64_t cap, int64_t* n);
int ssb_load_history(ssb_engine* e, const ssb_raw_record* rec, const ssb_sent_record* sent, int64_t n);
/* --diff_repeat, state change (SC) what-if query (runner.hpp:217-245: init_state(new state, (time, 0)), then the LP's
 * recorded events and sent events from (time, 0) on are re-armed): after ssb_load_history, LP `lp_id` (owned by this
 * rank) gets the state `words` and its recorded history with key >= (time, 0) becomes invalid — the repeat run
 * re-executes it with the new state and cancels what it had sent. */
int ssb_set_state_from(ssb_engine* e, int64_t lp_id, int64_t time, const uint32_t* words);

/* Back to the state right after ssb_load_states: empty calendar, zero statistics; model data, partition and the
 * loaded initial states are kept. Lets one engine run many simulations (bench steps) without re-allocating. */
int ssb_reset(ssb_engine* e);

/* change ssb_config.flags after creation (e.g. switch SSB_FLAG_KERNEL_TIMING on for one profiled run) */
int ssb_set_flags(ssb_engine* e, int32_t flags);

/* With SSB_FLAG_KERNEL_TIMING: per-kernel device time since create/reset. names[i] points to a static string.
 * Returns the number of kernels written (<= cap). */
int ssb_kernel_times(ssb_engine* e, const char** names, double* total_ms, int64_t* launches, int cap);

/* ---- multi-GPU: one engine per process/GPU ---- */
/* NCCL exchange mode: 128-byte NCCL unique id; created on rank 0 and distributed by the caller (torch.distributed, MPI, files, ...) */
int ssb_comm_unique_id(void* id128);
int ssb_comm_init(ssb_engine* e, const void* id128);
/* Peer-memory exchange mode (instead of ssb_comm_unique_id / ssb_comm_init; no NCCL involved): every engine
 * exports a 64-byte CUDA IPC handle of its mailbox; the caller gathers the handles of all ranks (rank order) and hands
 * the world x 64 bytes to every engine. The emit kernel then appends cross-partition events straight into the
 * destination GPU's mailbox over NVLink; at the start of every round the partitions write their GVT candidates into
 * each other's mailboxes and wait for the peers' (k_gvt) — the GVT min-reduction and the barrier in one step, without
 * a host-launched collective, so the whole window loop runs as one CUDA graph per GPU. All ranks must make the same
 * choice. One process per GPU (the exchange waits on the peers; ranks sharing a GPU could deadlock). */
int ssb_comm_export(ssb_engine* e, void* handle64);
int ssb_comm_import(ssb_engine* e, const void* handles, int32_t world);
/* Peer-memory exchange, ssb_run (graph loop) only: `windows` local windows per global synchronisation (default 1).
 * Between two synchronisations a partition keeps executing — every window covers the buckets from GVT up to one
 * bucket beyond its progress so far (at most 4 x `windows` buckets ahead of GVT) — while what it sends to peers
 * accumulates in their mailboxes; the peers integrate it at the next synchronisation, and
 * whatever arrives behind a partition's progress is a straggler there: the ordinary rollback machinery repairs it
 * (optimism ACROSS partitions, conservative windows inside one). Pays when few events cross partitions (TrafficSim
 * stripes: < 1 %) and a window is short (the GVT exchange then dominates); results do not depend on it. 1..64. */
int ssb_set_sync_interval(ssb_engine* e, int32_t windows);

/* ---- host helpers of the PHOLD app mirror (no device work) ---- */
/* phold.hpp:144-164: the two 10M-entry tables, libstdc++ default_random_engine streams */
int ssb_phold_build_tables(int64_t seed, double lambda, double remote_ratio, int64_t n,
                           int64_t* latency, int32_t* remote);
/* phold.hpp:197-209: initial events with id % stride == offset (id in [0, num_init_msg)) */
int64_t ssb_phold_init_events(int64_t num_lp, int64_t num_init_msg, const int64_t* latency, int64_t table_size,
                              int64_t offset, int64_t stride, ssb_event* out, int64_t cap);

/* ---- scenario ingestion at scale (no device work except ssb_scenario_load) ----
 * The reference reads its input as text, one getline + boost::split + atol per line and one shared_ptr object per
 * event / state (traffic_sim.hpp:345-451); a 10 M-trip scenario is ~3 GB of CSV. Two pieces replace that:
 *
 * 1. ssb_traffic_read: the three TrafficSim readers natively — the files are mapped, cut into line-aligned pieces and
 *    parsed by `threads` host threads (0 = all cores; at least ~1 MiB of text per thread; a negative value forces
 *    exactly that many pieces whatever the size) into flat arrays. Field meaning exactly as the reference's:
 *      partition file  traffic_reader::graph_read (traffic_sim.hpp:345-362): atol(line), line index = LP id
 *      road file       traffic_reader::road_read  (:364-408): one line per LP, roads split on ';', fields on ',':
 *                      state id = field 2 of the LAST road of the line, destination = field 3, speed = field 4,
 *                      length = field 5 (0 becomes 1), lanes = field 6 (unused by the handler)
 *      trip file       traffic_reader::trip_read  (:410-451): vehicle id = field 1, arrival = departure = field 3,
 *                      route = fields 4.., source = -1
 *    Numbers are read as C atol reads them (leading blanks, sign, digits, anything else ends the number). Lines end
 *    at '\n' (std::getline). A road with fewer than 7 fields or a trip with fewer than 5 is skipped (the reference
 *    indexes past the end of its token vector there). A file that cannot be opened is SSB_ERR_INVALID (the
 *    reference silently simulates nothing, traffic_sim.hpp:348-352). ssb_traffic_free releases the arrays.
 *
 * 2. The scenario image ("SSBSCN01"): everything an engine needs to start — partition table, LP states as the
 *    device's state words, model data slots, initial events as 32-byte device records — in one file whose sections
 *    are page-aligned, so opening it is one mmap and loading it is a handful of host-to-device copies straight from
 *    the mapping (no parse, no per-event object). ssb_traffic_compile writes one from the three CSV files;
 *    ssb_scenario_save writes one from arrays the caller holds (the drop-in runner does, after the application's own
 *    readers / generators ran once: SCALESIM_B200_SCENARIO_OUT); ssb_scenario_load puts an image into an engine of
 *    any rank / world: partition numbers are reduced modulo `world`, the states and events of LPs this rank owns are
 *    loaded (what runner::init does with the app's vectors, runner.hpp:98-147).
 */
typedef struct ssb_traffic_tables {
  int64_t n_part;   int64_t* part;                                         /* partition file (n_part = 0: none given) */
  int64_t n_states; int64_t* state_ids; int64_t* road_start;               /* road file: CSR, road_start[n_states + 1] */
  int64_t n_roads;  int64_t* road_dst; int64_t* road_speed; int64_t* road_len;
  int64_t n_trips;  int64_t* trip_id; int64_t* trip_dep; int64_t* route_start;   /* trip file: CSR, route_start[n_trips + 1] */
  int64_t n_nodes;  int64_t* route_nodes;
} ssb_traffic_tables;
int ssb_traffic_read(const char* partition_file /* may be NULL */, const char* road_file /* may be NULL */,
                     const char* trip_file /* may be NULL */, int threads, ssb_traffic_tables* out);
void ssb_traffic_free(ssb_traffic_tables* t);
/* the inverse, for feeding a generated scenario to the reference's own binary: files in the reference's format
 * (road line: "R,{road id},{lp},{dst},{speed},{length},1" joined by ';'; trip line: "{i},{id},0,{departure},{nodes...}") */
int ssb_traffic_write_csv(const ssb_traffic_tables* t, const char* partition_file /* may be NULL */,
                          const char* road_file /* may be NULL */, const char* trip_file /* may be NULL */);

#define SSB_SCENARIO_SLOTS 4
typedef struct ssb_scenario {
  int32_t model;                      /* ssb_model_id */
  int32_t state_words;                /* 32-bit words per LP state (= ssb_state_words of an engine of that model) */
  int64_t n_lp, finish_time;
  int64_t n_part;   const int64_t* lp_to_part;     /* partition numbers as the application gave them (0: lp % world) */
  int64_t n_states; const int64_t* state_lps; const uint32_t* states;      /* states[n_states * state_words] */
  int64_t slot_bytes[SSB_SCENARIO_SLOTS]; const void* slot_data[SSB_SCENARIO_SLOTS];   /* ssb_set_model_data slots */
  int64_t n_events; const ssb_raw_record* events;  /* initial events, every partition's */
  void* mapping; int64_t mapping_bytes;            /* set by ssb_scenario_open (the pointers above point into it) */
} ssb_scenario;
int ssb_scenario_save(const char* path, const ssb_scenario* sc);
int ssb_scenario_open(const char* path, ssb_scenario* sc);      /* read-only mapping; ssb_scenario_close unmaps */
void ssb_scenario_close(ssb_scenario* sc);
/* partition + model data + the states and events of this engine's rank; call it where ssb_set_partition would be */
int ssb_scenario_load(ssb_engine* e, const ssb_scenario* sc);
/* CSV -> image for TrafficSim: ssb_traffic_read + the packing of scalesim/models/traffic_model.hpp (state words, road
 * table for junctions with more than two out-roads, route pool, one initial event per trip) + ssb_scenario_save */
int ssb_traffic_compile(const char* partition_file /* may be NULL */, const char* road_file, const char* trip_file,
                        int64_t finish_time, int threads, const char* image_path);
/* the configuration an engine was created with */
int ssb_get_config(ssb_engine* e, ssb_config* out);

#ifdef __cplusplus
}
#endif
#endif /* SCALESIM_B200_H_ */
