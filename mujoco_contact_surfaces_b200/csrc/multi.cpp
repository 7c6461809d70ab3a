// hcs_multi: one context spanning several GPUs (include/hcs.h "one context spanning several GPUs").
//
// The reference runs the path once per mjData on the one physics thread (mujoco_contact_surfaces_plugin.cpp:88-95); the
// north star batches independent environments and shards them by index over the GPUs of a box.  This layer does the
// sharding INSIDE the library, in C++: D single-device contexts (blocks), each with a contiguous range of environments
// and a worker thread that issues the block's calls with the block's device current.  Configuration calls are replayed
// on every block; a step gives every block its slice of the caller's env-major arrays.  No collective: environments
// exchange nothing.  Host code only (the blocks are driven through the C ABI), so it carries no device code of its own.
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hcs.h"

namespace {

// one worker per block: runs the jobs it is handed, in order.  post() never blocks; wait() returns when everything posted so
// far has run, with the first negative return code since the last wait() (0 if none).  Both sides spin for a short while
// before they sleep on the condition variable: a step is ~0.1 ms, and two sleeping hand-offs per step (caller -> worker ->
// caller, 5 - 20 us each) showed in the 8-GPU rate.
class Worker {
public:
	Worker() : th_([this] { loop(); }) {}
	~Worker()
	{
		{
			std::lock_guard<std::mutex> l(mu_);
			quit_ = true;
		}
		cv_.notify_all();
		th_.join();
	}
	void post(std::function<int()> job)
	{
		{
			std::lock_guard<std::mutex> l(mu_);
			jobs_.push_back(std::move(job));
			posted_.fetch_add(1, std::memory_order_release);
		}
		cv_.notify_all();
	}
	int wait()
	{
		const long target = posted_.load(std::memory_order_acquire);
		for (int spin = 0; spin < SPIN && done_.load(std::memory_order_acquire) < target; ++spin)
			cpu_relax();
		std::unique_lock<std::mutex> l(mu_);
		cv_.wait(l, [this, target] { return done_.load(std::memory_order_acquire) >= target; });
		const int rc = first_error_ < 0 ? first_error_ : last_rc_;
		first_error_ = 0;
		return rc;
	}

private:
	static constexpr int SPIN = 4000; // ~0.1 - 0.2 ms of pause instructions
	static void cpu_relax()
	{
#if defined(__x86_64__) || defined(__i386__)
		__builtin_ia32_pause();
#endif
	}
	void loop()
	{
		for (;;) {
			std::function<int()> job;
			{
				for (int spin = 0; spin < SPIN && posted_.load(std::memory_order_acquire) == taken_; ++spin)
					cpu_relax();
				std::unique_lock<std::mutex> l(mu_);
				cv_.wait(l, [this] { return !jobs_.empty() || quit_; });
				if (jobs_.empty())
					return;
				job = std::move(jobs_.front());
				jobs_.pop_front();
				++taken_;
			}
			const int rc = job();
			{
				std::lock_guard<std::mutex> l(mu_);
				last_rc_ = rc;
				if (rc < 0 && first_error_ == 0)
					first_error_ = rc;
				done_.fetch_add(1, std::memory_order_release);
			}
			cv_.notify_all();
		}
	}
	std::mutex mu_;
	std::condition_variable cv_;
	std::deque<std::function<int()>> jobs_;
	std::atomic<long> posted_{ 0 }, done_{ 0 };
	long taken_ = 0; // (worker thread only)
	bool quit_   = false;
	int last_rc_ = 0, first_error_ = 0;
	std::thread th_; // last member: the thread starts when everything above exists
};

struct Block {
	hcs_ctx *ctx = nullptr;
	int device = 0, start = 0, count = 0;
	Worker *worker = nullptr;
	int64_t ticket = -1;                     // the block's own ticket of the last pipelined step
	std::vector<float *> img, curved, taxel; // per-block views of the caller's output arrays (pipelined step)
};

} // namespace

struct hcs_multi {
	std::vector<Block> blocks;
	int n_envs = 0;
	std::string err;
	int64_t next_ticket = 0;
	std::vector<std::pair<int, int>> sensor_dims;
};

static thread_local std::string g_multi_create_error;

// runs f(block) on every block's worker concurrently; first error wins (its text is kept)
static int fan_out(hcs_multi *m, const std::function<int(Block &)> &f)
{
	for (Block &b : m->blocks)
		if (b.count > 0)
			b.worker->post([&f, &b] { return f(b); });
	int rc = HCS_OK;
	for (Block &b : m->blocks) {
		if (b.count == 0)
			continue;
		int r = b.worker->wait();
		if (r < 0 && rc >= 0) {
			rc     = r;
			m->err = std::string("block on device ") + std::to_string(b.device) + ": " + hcs_last_error(b.ctx);
		} else if (rc >= 0) {
			rc = r; // configuration indices: identical on every block
		}
	}
	return rc;
}

extern "C" {

int hcs_multi_create(const hcs_config *cfg, const int *devices, int n_devices, hcs_multi **out)
{
	if (!cfg || !devices || !out || n_devices < 1 || n_devices > 64 || cfg->n_envs < 1) {
		g_multi_create_error = "hcs_multi_create: needs a config with n_envs >= 1 and 1 .. 64 devices";
		return HCS_E_INVALID;
	}
	hcs_multi *m = new hcs_multi();
	m->n_envs    = cfg->n_envs;
	const int base = cfg->n_envs / n_devices, rem = cfg->n_envs % n_devices;
	for (int k = 0; k < n_devices; ++k) {
		Block b;
		b.device = devices[k];
		b.start  = k * base + (k < rem ? k : rem);
		b.count  = base + (k < rem ? 1 : 0);
		m->blocks.push_back(b);
	}
	int rc = HCS_OK;
	for (Block &b : m->blocks) {
		if (b.count == 0)
			continue;
		b.worker       = new Worker();
		hcs_config c   = *cfg;
		c.device       = b.device;
		c.n_envs       = b.count;
		c.stream       = nullptr;
		if (c.max_tactile_triangles > 0) // the pool is sized for the whole batch: this block's share
			c.max_tactile_triangles = (int)(((long long)c.max_tactile_triangles * b.count + cfg->n_envs - 1) / cfg->n_envs);
		if (c.max_faces > 0)
			c.max_faces = (int)(((long long)c.max_faces * b.count + cfg->n_envs - 1) / cfg->n_envs);
		hcs_ctx **slot = &b.ctx;
		b.worker->post([c, slot] { return hcs_create(&c, slot); });
		int r = b.worker->wait();
		if (r < 0 && rc >= 0) {
			rc                   = r;
			g_multi_create_error = std::string("device ") + std::to_string(b.device) + ": " + hcs_last_error(nullptr);
		}
	}
	if (rc < 0) {
		hcs_multi_destroy(m);
		return rc;
	}
	*out = m;
	return HCS_OK;
}

void hcs_multi_destroy(hcs_multi *m)
{
	if (!m)
		return;
	for (Block &b : m->blocks) {
		if (b.worker && b.ctx) {
			hcs_ctx *c = b.ctx;
			b.worker->post([c] {
				hcs_destroy(c);
				return 0;
			});
			b.worker->wait();
		}
		delete b.worker;
	}
	delete m;
}

const char *hcs_multi_last_error(const hcs_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int hcs_multi_n_blocks(const hcs_multi *m) { return m ? (int)m->blocks.size() : HCS_E_INVALID; }

int hcs_multi_block(const hcs_multi *m, int k, int *env_start, int *env_count, hcs_ctx **ctx)
{
	if (!m || k < 0 || k >= (int)m->blocks.size())
		return HCS_E_INVALID;
	if (env_start)
		*env_start = m->blocks[k].start;
	if (env_count)
		*env_count = m->blocks[k].count;
	if (ctx)
		*ctx = m->blocks[k].ctx;
	return HCS_OK;
}

int hcs_multi_add_geom(hcs_multi *m, int type, const double size[3], const float *mesh_vert, int n_vert, const int32_t *mesh_face,
                       int n_face, const double props[5])
{
	if (!m)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_add_geom(b.ctx, type, size, mesh_vert, n_vert, mesh_face, n_face, props); });
}

int hcs_multi_add_soft_mesh(hcs_multi *m, const double *verts, int n_vert, const int32_t *tets, int n_tet,
                            const double *vertex_pressure, const double props[5])
{
	if (!m)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_add_soft_mesh(b.ctx, verts, n_vert, tets, n_tet, vertex_pressure, props); });
}

int hcs_multi_add_rigid_mesh(hcs_multi *m, const double *verts, int n_vert, const int32_t *tris, int n_tri, const double props[5])
{
	if (!m)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_add_rigid_mesh(b.ctx, verts, n_vert, tris, n_tri, props); });
}

int hcs_multi_update_geom(hcs_multi *m, int geom, const double size[3])
{
	if (!m)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_update_geom(b.ctx, geom, size); });
}

int hcs_multi_set_env_sizes(hcs_multi *m, int geom, const double *sizes)
{
	if (!m)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_set_env_sizes(b.ctx, geom, sizes ? sizes + 3 * (size_t)b.start : nullptr); });
}

int hcs_multi_set_pairs(hcs_multi *m, const int32_t *g1, const int32_t *g2, int n_pairs)
{
	if (!m)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_set_pairs(b.ctx, g1, g2, n_pairs); });
}

int hcs_multi_add_flat_sensor(hcs_multi *m, int geom, double resolution, int sampling_resolution, int window, float sigma)
{
	if (!m)
		return HCS_E_INVALID;
	int s = fan_out(m, [=](Block &b) { return hcs_add_flat_sensor(b.ctx, geom, resolution, sampling_resolution, window, sigma); });
	if (s >= 0) {
		int cx = 0, cy = 0;
		for (Block &b : m->blocks)
			if (b.count > 0) {
				hcs_sensor_dims(b.ctx, s, &cx, &cy);
				break;
			}
		if ((int)m->sensor_dims.size() <= s)
			m->sensor_dims.resize(s + 1);
		m->sensor_dims[s] = { cx, cy };
	}
	return s;
}

int hcs_multi_finalize(hcs_multi *m)
{
	if (!m)
		return HCS_E_INVALID;
	return fan_out(m, [](Block &b) { return hcs_finalize(b.ctx); });
}

int hcs_multi_step(hcs_multi *m, const double *xpos, const double *xmat, const double *vel, int with_sensors)
{
	if (!m || !xpos || !xmat || !vel)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) {
		const size_t ng = (size_t)hcs_n_geoms(b.ctx), o = (size_t)b.start * ng;
		return hcs_step(b.ctx, xpos + o * 3, xmat + o * 9, vel + o * 6, with_sensors);
	});
}

int hcs_multi_step_async(hcs_multi *m, const double *xpos, const double *xmat, const double *vel, int with_sensors,
                         const hcs_outputs *out, int64_t *ticket)
{
	if (!m || !xpos || !xmat || !vel || !ticket)
		return HCS_E_INVALID;
	// The call only hands the step to the blocks' workers and returns: every worker queues its slice through the block's
	// pipelined entry point on its own time, so the caller's thread costs a few microseconds per step however many GPUs
	// there are (waiting for all enqueues cost 0.116 ms per step on 8 GPUs against 0.091 ms of GPU time).  Errors of the
	// enqueue are reported by hcs_multi_wait.  The caller's arrays must stay untouched until hcs_multi_wait(ticket).
	const size_t n_img = m->sensor_dims.size();
	hcs_outputs o{};
	std::vector<float *> images;
	if (out) {
		o = *out;
		if (out->sensor_images && with_sensors)
			images.assign(out->sensor_images, out->sensor_images + n_img);
	}
	for (Block &b : m->blocks) {
		if (b.count == 0)
			continue;
		Block *bp = &b;
		b.worker->post([=]() {
			Block &blk      = *bp;
			const size_t ng = (size_t)hcs_n_geoms(blk.ctx), np = (size_t)hcs_n_pairs(blk.ctx), off = (size_t)blk.start * ng;
			hcs_outputs mine{};
			mine.geom_wrench  = o.geom_wrench ? o.geom_wrench + off * 6 : nullptr;
			mine.pair_results = o.pair_results ? o.pair_results + (size_t)blk.start * np : nullptr;
			if (!images.empty()) {
				blk.img.assign(n_img, nullptr);
				for (size_t s = 0; s < n_img; ++s)
					if (images[s])
						blk.img[s] = images[s] + (size_t)blk.start * m->sensor_dims[s].first * m->sensor_dims[s].second;
				mine.sensor_images = blk.img.data();
			}
			// curved / taxel sensors are configured per block (hcs_multi_block): their outputs are not mirrored here
			return hcs_step_async(blk.ctx, xpos + off * 3, xmat + off * 9, vel + off * 6, with_sensors, &mine, &blk.ticket);
		});
	}
	*ticket = m->next_ticket++;
	return HCS_OK;
}

int hcs_multi_wait(hcs_multi *m, int64_t ticket)
{
	if (!m || ticket < 0 || ticket >= m->next_ticket)
		return HCS_E_INVALID;
	// every block has taken part in every pipelined step, so its own ticket numbers run in step with ours
	const int64_t behind = m->next_ticket - 1 - ticket;
	return fan_out(m, [=](Block &b) { return hcs_wait(b.ctx, b.ticket - behind); });
}

int hcs_multi_get_geom_wrenches(hcs_multi *m, double *out)
{
	if (!m || !out)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_get_geom_wrenches(b.ctx, out + (size_t)b.start * hcs_n_geoms(b.ctx) * 6); });
}

int hcs_multi_get_pair_results(hcs_multi *m, hcs_pair_result *out)
{
	if (!m || !out)
		return HCS_E_INVALID;
	return fan_out(m, [=](Block &b) { return hcs_get_pair_results(b.ctx, out + (size_t)b.start * hcs_n_pairs(b.ctx)); });
}

int hcs_multi_get_sensor_image(hcs_multi *m, int sensor, float *out)
{
	if (!m || !out || sensor < 0 || sensor >= (int)m->sensor_dims.size())
		return HCS_E_INVALID;
	const size_t per = (size_t)m->sensor_dims[sensor].first * m->sensor_dims[sensor].second;
	return fan_out(m, [=](Block &b) { return hcs_get_sensor_image(b.ctx, sensor, out + (size_t)b.start * per); });
}

} // extern "C"
