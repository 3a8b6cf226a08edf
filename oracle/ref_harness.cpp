// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Thin C-callable harness around the UNMODIFIED reference CPU build
// (manasakani/DeviceKMC, compiled from /root/reference/src by oracle/Makefile into
// oracle/_ref/).  It replaces the reference's kmc_main.cpp (which cannot run in a
// CPU-only build: CreateCublasHandle has no return statement, utils.cpp:368-379)
// and drives the reference's own Device / KMCProcess objects:
//
//   Device::Device                      Device.cpp:17
//   Device::makeSubstoichiometric       Device.cpp:202
//   Device::updateCharge                potential_solver.cpp:142
//   Device::background_potential        potential_solver.cpp:289   (dense K + dgesv)
//   Device::poisson_gridless            potential_solver.cpp:412
//   KMCProcess::update_events_and_rates KMCProcess.cpp:67
//   KMCProcess::executeKMCStep          KMCProcess.cpp:259
//
// This translation unit is compiled with -fno-access-control so the private
// members (background_potential, poisson_gridless, random_generator) can be
// reached without touching the reference sources.
#include "Device.h"
#include "KMCProcess.h"
#include "input_parser.h"

#include <cstring>
#include <memory>

namespace {
struct RefSim {
    std::unique_ptr<KMCParameters> p;
    std::unique_ptr<Device> dev;
    std::unique_ptr<KMCProcess> sim;
    GPUBuffers gpubuf;   // empty (CPU-only build)
    int last_num_events = 0;
};
}  // namespace

extern "C" {

// params: a reference-format parameters.txt.  xyz_a / xyz_b: structure files
// (xyz_b may be NULL or "").  If `substoich` != 0 and p.pristine, the initial
// vacancies are drawn exactly as kmc_main.cpp:102-103 does.
void *ref_create(const char *params, const char *xyz_a, const char *xyz_b, int substoich) {
    auto *s = new RefSim;
    s->p.reset(new KMCParameters(std::string(params)));
    std::vector<std::string> files;
    files.push_back(xyz_a);
    if (xyz_b && xyz_b[0]) files.push_back(xyz_b);
    s->dev.reset(new Device(files, *s->p));
    if (substoich && s->p->pristine) s->dev->makeSubstoichiometric(s->p->initial_vacancy_concentration);
    s->sim.reset(new KMCProcess(*s->dev, s->p->freq));
    return s;
}

void ref_destroy(void *h) { delete static_cast<RefSim *>(h); }

// out[0..11]: N, nn, N_atom, num_atoms_contact, num_atoms_first_layer, pbc, num_metals, rnd_seed
void ref_info_int(void *h, int *out) {
    auto *s = static_cast<RefSim *>(h);
    out[0] = s->dev->N;
    out[1] = s->dev->max_num_neighbors;
    out[2] = s->dev->N_atom;
    out[3] = s->p->num_atoms_contact;
    out[4] = s->p->num_atoms_first_layer;
    out[5] = s->dev->pbc ? 1 : 0;
    out[6] = (int)s->p->metals.size();
    out[7] = (int)s->p->rnd_seed;
}

// out: sigma, k, T_bg, freq, nn_dist, high_G, low_G, lattice[0..2]
void ref_info_double(void *h, double *out) {
    auto *s = static_cast<RefSim *>(h);
    out[0] = s->dev->sigma;
    out[1] = s->dev->k;
    out[2] = s->dev->T_bg;
    out[3] = s->sim->freq;
    out[4] = s->dev->nn_dist;
    out[5] = s->p->high_G;
    out[6] = s->p->low_G;
    out[7] = s->dev->lattice[0];
    out[8] = s->dev->lattice[1];
    out[9] = s->dev->lattice[2];
}

void ref_get_metals(void *h, int *out) {
    auto *s = static_cast<RefSim *>(h);
    for (size_t i = 0; i < s->p->metals.size(); ++i) out[i] = (int)s->p->metals[i];
}

// layer table used by the rate table: out[5*l + {0..3}] = E_gen, E_rec, E_Vdiff, E_Odiff; returns #layers
int ref_get_layers(void *h, double *out) {
    auto *s = static_cast<RefSim *>(h);
    int l = 0;
    for (auto &L : s->sim->layers) {
        out[4 * l + 0] = L.E_gen_0;
        out[4 * l + 1] = L.E_rec_1;
        out[4 * l + 2] = L.E_diff_2;
        out[4 * l + 3] = L.E_diff_3;
        ++l;
    }
    return l;
}

void ref_get_positions(void *h, double *x, double *y, double *z) {
    auto *s = static_cast<RefSim *>(h);
    size_t n = s->dev->N;
    std::memcpy(x, s->dev->site_x.data(), n * sizeof(double));
    std::memcpy(y, s->dev->site_y.data(), n * sizeof(double));
    std::memcpy(z, s->dev->site_z.data(), n * sizeof(double));
}

void ref_get_neigh_idx(void *h, int *out) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(out, s->dev->neigh_idx.data(), s->dev->neigh_idx.size() * sizeof(int));
}

void ref_get_site_layer(void *h, int *out) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(out, s->sim->site_layer.data(), s->sim->site_layer.size() * sizeof(int));
}

void ref_get_element(void *h, int *out) {
    auto *s = static_cast<RefSim *>(h);
    for (int i = 0; i < s->dev->N; ++i) out[i] = (int)s->dev->site_element[i];
}
void ref_set_element(void *h, const int *in) {
    auto *s = static_cast<RefSim *>(h);
    for (int i = 0; i < s->dev->N; ++i) s->dev->site_element[i] = (ELEMENT)in[i];
}
void ref_get_charge(void *h, int *out) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(out, s->dev->site_charge.data(), s->dev->N * sizeof(int));
}
void ref_set_charge(void *h, const int *in) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(s->dev->site_charge.data(), in, s->dev->N * sizeof(int));
}
void ref_get_potential_boundary(void *h, double *out) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(out, s->dev->site_potential_boundary.data(), s->dev->N * sizeof(double));
}
void ref_set_potential_boundary(void *h, const double *in) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(s->dev->site_potential_boundary.data(), in, s->dev->N * sizeof(double));
}
void ref_get_potential_charge(void *h, double *out) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(out, s->dev->site_potential_charge.data(), s->dev->N * sizeof(double));
}
void ref_set_potential_charge(void *h, const double *in) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(s->dev->site_potential_charge.data(), in, s->dev->N * sizeof(double));
}

// Device::updateCharge, CPU branch (potential_solver.cpp:172-217)
void ref_update_charge(void *h) {
    auto *s = static_cast<RefSim *>(h);
    s->dev->updateCharge(s->gpubuf, s->p->metals);
}

// Device::background_potential with the contact size the CPU path uses
// (num_atoms_contact, potential_solver.cpp:271,294).  `n_contact` <= 0 keeps it.
void ref_background_potential(void *h, double Vd, int n_contact) {
    auto *s = static_cast<RefSim *>(h);
    int nc = n_contact > 0 ? n_contact : s->p->num_atoms_contact;
    s->dev->background_potential(nullptr, nc, Vd, s->dev->lattice, s->p->G_coeff, s->p->high_G,
                                 s->p->low_G, s->p->metals, 0);
}

// Device::setLaplacePotential, CPU branch (potential_solver.cpp:4-139): fills site_CB_edge.
// It sizes the contacts by num_atoms_first_layer (:7-8).
void ref_laplace_cb_edge(void *h, double Vd) {
    auto *s = static_cast<RefSim *>(h);
    s->dev->setLaplacePotential(nullptr, nullptr, s->gpubuf, *s->p, Vd);
}
void ref_get_cb_edge(void *h, double *out) {
    auto *s = static_cast<RefSim *>(h);
    std::memcpy(out, s->dev->site_CB_edge.data(), s->dev->N * sizeof(double));
}

// Device::writeSnapshot (Device.cpp:236-252): writes ./<folder>/<filename>
void ref_write_snapshot(void *h, const char *filename, const char *folder) {
    auto *s = static_cast<RefSim *>(h);
    s->dev->writeSnapshot(filename, folder);
}

// Device::poisson_gridless (potential_solver.cpp:412-432)
void ref_poisson_gridless(void *h) {
    auto *s = static_cast<RefSim *>(h);
    s->dev->poisson_gridless(s->p->num_atoms_contact, s->dev->lattice);
}

// KMCProcess::update_events_and_rates (KMCProcess.cpp:67-164)
void ref_rate_table(void *h, int *event_type, double *event_prob) {
    auto *s = static_cast<RefSim *>(h);
    size_t n = (size_t)s->dev->N * s->dev->max_num_neighbors;
    std::vector<EVENTTYPE> et(n);
    s->sim->update_events_and_rates(*s->dev, et.data(), event_prob);
    for (size_t i = 0; i < n; ++i) event_type[i] = (int)et[i];
}

// KMCProcess::executeKMCStep, CPU branch (KMCProcess.cpp:282-365).  Returns the number of
// executed events; the (i, j) pair of each executed event, in order, is decoded from
// KMCProcess::affected_neighborhood (KMCProcess.cpp:166-185 pushes 2*nn slot indices and
// then every conflicting table index for each event).
int ref_kmc_step(void *h, double *step_time, int *events_ij, int max_events) {
    auto *s = static_cast<RefSim *>(h);
    s->sim->executeKMCStep(s->gpubuf, *s->dev, step_time);
    const std::vector<int> &a = s->sim->affected_neighborhood;
    const int nn = s->dev->max_num_neighbors;
    const long total = (long)s->dev->N * nn;
    size_t pos = 0;
    int ne = 0;
    while (pos < a.size()) {
        int i = a[pos] / nn;
        int j = a[pos + 1] / nn;
        pos += 2 * (size_t)nn;
        long cnt = 0;
        for (long idx = 0; idx < total; ++idx) {
            int i_ = (int)(idx / nn);
            int j_ = s->dev->neigh_idx[idx];
            if (i == i_ || j == j_ || i == j_ || j == i_) ++cnt;
        }
        pos += (size_t)cnt;
        if (ne < max_events) {
            events_ij[2 * ne] = i;
            events_ij[2 * ne + 1] = j;
        }
        ++ne;
    }
    s->last_num_events = ne;
    return ne;
}

// peek at the next `n` doubles of the KMC random stream without consuming it
// (RandomNumberGenerator, random_num.h:4-23; seed rnd_seed_kmc, structure_input.h:8)
void ref_peek_kmc_rng(void *h, double *out, int n) {
    auto *s = static_cast<RefSim *>(h);
    RandomNumberGenerator copy = s->sim->random_generator;
    for (int i = 0; i < n; ++i) out[i] = copy.getRandomNumber();
}

// the reference's site_dist (utils.cpp:100-137) for spot checks
double ref_site_dist(double x1, double y1, double z1, double x2, double y2, double z2,
                     const double *lattice, int pbc) {
    std::vector<double> lat(lattice, lattice + 3);
    return site_dist(x1, y1, z1, x2, y2, z2, lat, pbc != 0);
}

}  // extern "C"
