// matrix_free_adaptor.h -- builds the device mirror of dealii::MatrixFree (adaflo_b200::MatrixFreeData) from the data
// MatrixFree::reinit produced on the host: INTEGRATION.md section 2 as code.  The adaptor reads deal.II through a small
// view (MatrixFreeView below) whose members name the deal.II accessors they stand for; where deal.II exists the view is
// a dozen forwarding lambdas around dealii::MatrixFree<dim> (INTEGRATION.md shows them), tests/cpp/adaptor_driver.cpp
// fills it from a mock of the batch / lane (SoA over VectorizedArray lanes) storage.
//
// What it handles:
//  * de-interleaving of the cell batches (n_active_entries_per_cell_batch lanes per batch),
//  * ANY dof numbering of a vector-valued space: deal.II's plain_dof_indices are component-blocked per cell
//    (idx[c * nodes_per_cell + node]); the library wants node-major vectors, so the adaptor numbers the nodes itself
//    and records where each library entry lives in the host vector (adaflo_b200_mf_set_dof_numbering).  This covers
//    DoFRenumbering::Cuthill_McKee on the velocity DoFHandler (source/navier_stokes.cc:208-209),
//  * constrained dofs (homogeneous Dirichlet) -> flag bits of the cell lists and the constrained-row list,
//  * Cartesian / affine cells: one table row per distinct Jacobian (deal.II's compression), general cells: per point.
// Ghost exchange plans (MatrixFree::get_vector_partitioner) are passed through unchanged.
#pragma once

#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

#include "adaflo_b200/navier_stokes_matrix.h"

namespace adaflo_b200
{
  /// one DoFHandler of the MatrixFree object
  struct DoFHandlerView
  {
    int          fe_kind        = ADAFLO_B200_FE_Q;  ///< FE_Q / FE_Q_DG0 / FE_Q_iso_Q1
    int          degree         = 1;
    unsigned int n_components   = 1;
    unsigned int nodes_per_cell = 8;   ///< scalar dofs per cell and component (lexicographic, x fastest; DG0 dof last)
    std::uint64_t n_owned_dofs  = 0;   ///< locally_owned_size() of the vectors of this DoFHandler
    std::uint64_t n_ghost_dofs  = 0;
    /// MatrixFree::get_dof_info(dof_index).plain_dof_indices of cell (batch, lane): n_components * nodes_per_cell local
    /// vector indices, component-blocked, owned dofs first (index < n_owned_dofs), then ghosts
    std::function<const unsigned int *(unsigned int batch, unsigned int lane)> plain_dof_indices;
    /// MatrixFree::get_constrained_dofs(dof_index): local indices of the (homogeneously) constrained owned dofs, sorted
    std::vector<unsigned int> constrained_dofs;
    /// per-cell refinement configuration of cells with hanging faces / edges or empty
    std::function<std::uint32_t(unsigned int batch, unsigned int lane)> hanging_mask;
    std::vector<adaflo_b200_exchange> exchange;  ///< in terms of LIBRARY node indices (see node_of_dof after build)
  };

  /// one quadrature formula (QGauss(n_q_1d) or QIterated(QGauss(2), n_sub)) with the mapping data of MappingInfo
  struct QuadratureView
  {
    int n_q_1d = 2;
    /// mapping_info.cell_type[batch]: 0 cartesian, 1 affine, 2 general (the maximum over the batches decides the layout)
    std::function<int(unsigned int batch)> cell_type;
    /// inverse Jacobian d xi_f / d x_e (entry 3 f + e) and |J| of cell (batch, lane); q ignored unless general
    std::function<void(unsigned int batch, unsigned int lane, unsigned int q, double (&inv_jac)[9], double &det_or_jxw)> geometry;
  };

  struct MatrixFreeView
  {
    unsigned int                                     n_cell_batches = 0;
    std::function<unsigned int(unsigned int batch)>  n_active_entries_per_cell_batch;
    std::vector<DoFHandlerView>                      dof_handlers;
    std::vector<QuadratureView>                      quadratures;
  };

  /// result: the device mirror plus the maps between deal.II's numbering and the library's
  struct AdaptedMatrixFree
  {
    std::unique_ptr<MatrixFreeData>         matrix_free;
    std::vector<std::vector<std::uint32_t>> host_position;  ///< per space: position of library entry i in the host vector
    std::vector<std::pair<unsigned int, unsigned int>> cell_of;  ///< library cell -> (batch, lane)
  };

  inline AdaptedMatrixFree build_matrix_free(const Context &ctx, const MatrixFreeView &view)
  {
    AdaptedMatrixFree out;
    for (unsigned int b = 0; b < view.n_cell_batches; ++b)
      for (unsigned int l = 0; l < view.n_active_entries_per_cell_batch(b); ++l) out.cell_of.emplace_back(b, l);
    const std::uint32_t n_cells = static_cast<std::uint32_t>(out.cell_of.size());

    const std::size_t n_spaces = view.dof_handlers.size();
    std::vector<std::vector<std::uint32_t>> cell_nodes(n_spaces), masks(n_spaces), constrained(n_spaces);
    std::vector<adaflo_b200_space_desc>     spaces(n_spaces);
    out.host_position.resize(n_spaces);
    for (std::size_t s = 0; s < n_spaces; ++s)
      {
        const DoFHandlerView &dh = view.dof_handlers[s];
        const unsigned int    nc = dh.n_components, npc = dh.nodes_per_cell;
        if (dh.n_owned_dofs % nc || dh.n_ghost_dofs % nc)
          throw std::invalid_argument("build_matrix_free: dof counts are not multiples of the component count");
        const std::uint32_t n_owned_nodes = static_cast<std::uint32_t>(dh.n_owned_dofs / nc), n_ghost_nodes = static_cast<std::uint32_t>(dh.n_ghost_dofs / nc);
        // (1) node numbering: a node is identified by the dof of its component 0; owned nodes in increasing order of that
        //     dof, ghost nodes in increasing order behind them (the order of the partitioner's ghost range)
        constexpr std::uint32_t none = 0xffffffffu;
        std::vector<std::uint32_t> node_of_dof0(dh.n_owned_dofs + dh.n_ghost_dofs, none);
        for (const auto &cell : out.cell_of)
          {
            const unsigned int *idx = dh.plain_dof_indices(cell.first, cell.second);
            for (unsigned int n = 0; n < npc; ++n) node_of_dof0[idx[n]] = 0;  // mark
          }
        std::uint32_t next_owned = 0, next_ghost = n_owned_nodes;
        for (std::uint64_t d = 0; d < node_of_dof0.size(); ++d)
          if (node_of_dof0[d] != none)
            node_of_dof0[d] = d < dh.n_owned_dofs ? next_owned++ : next_ghost++;
        if (next_owned > n_owned_nodes || next_ghost > n_owned_nodes + n_ghost_nodes)
          throw std::invalid_argument("build_matrix_free: more nodes than dofs / n_components");
        // owned nodes no cell of this rank touches keep their place at the end of the owned range
        std::vector<std::uint32_t> &pos = out.host_position[s];
        pos.assign(dh.n_owned_dofs, none);
        std::vector<unsigned char> is_constrained(dh.n_owned_dofs, 0);
        for (const unsigned int d : dh.constrained_dofs)
          if (d < dh.n_owned_dofs)
            is_constrained[d] = 1;
        // (2) cell lists with flag bits, host positions of the library entries
        cell_nodes[s].resize(static_cast<std::size_t>(n_cells) * npc);
        for (std::uint32_t c = 0; c < n_cells; ++c)
          {
            const unsigned int *idx = dh.plain_dof_indices(out.cell_of[c].first, out.cell_of[c].second);
            for (unsigned int n = 0; n < npc; ++n)
              {
                const std::uint32_t node  = node_of_dof0[idx[n]];
                std::uint32_t       entry = node;
                for (unsigned int comp = 0; comp < nc; ++comp)
                  {
                    const unsigned int d = idx[comp * npc + n];
                    if (d < dh.n_owned_dofs)
                      {
                        pos[static_cast<std::size_t>(node) * nc + comp] = d;
                        if (is_constrained[d])
                          entry |= 1u << (ADAFLO_B200_FLAG_SHIFT + comp);
                      }
                  }
                cell_nodes[s][static_cast<std::size_t>(c) * npc + n] = entry;
              }
          }
        {  // entries of untouched owned nodes: any free host positions, in order (keeps pos a permutation)
          std::vector<unsigned char> used(dh.n_owned_dofs, 0);
          for (const std::uint32_t p : pos)
            if (p != none)
              used[p] = 1;
          std::uint64_t free_pos = 0;
          for (std::uint32_t &p : pos)
            if (p == none)
              {
                while (used[free_pos]) ++free_pos;
                p             = static_cast<std::uint32_t>(free_pos);
                used[free_pos] = 1;
              }
        }
        // (3) constrained rows in library numbering
        std::vector<std::uint32_t> lib_of_host(dh.n_owned_dofs);
        for (std::uint64_t i = 0; i < pos.size(); ++i) lib_of_host[pos[i]] = static_cast<std::uint32_t>(i);
        for (const unsigned int d : dh.constrained_dofs)
          if (d < dh.n_owned_dofs)
            constrained[s].push_back(lib_of_host[d]);
        std::sort(constrained[s].begin(), constrained[s].end());
        if (dh.hanging_mask)
          {
            masks[s].resize(n_cells);
            for (std::uint32_t c = 0; c < n_cells; ++c) masks[s][c] = dh.hanging_mask(out.cell_of[c].first, out.cell_of[c].second);
          }
        adaflo_b200_space_desc &sd = spaces[s];
        sd                  = adaflo_b200_space_desc{};
        sd.fe_kind          = dh.fe_kind;
        sd.degree           = dh.degree;
        sd.n_components     = static_cast<int>(nc);
        sd.n_owned_nodes    = n_owned_nodes;
        sd.n_ghost_nodes    = n_ghost_nodes;
        sd.nodes_per_cell   = npc;
        sd.cell_nodes       = cell_nodes[s].data();
        sd.hanging_masks    = masks[s].empty() ? nullptr : masks[s].data();
        sd.n_constrained    = static_cast<std::uint32_t>(constrained[s].size());
        sd.constrained_dofs = constrained[s].data();
        sd.n_exchange       = static_cast<int>(dh.exchange.size());
        sd.exchange         = dh.exchange.empty() ? nullptr : dh.exchange.data();
      }

    // (4) geometry: one table row per distinct Jacobian for Cartesian / affine meshes, per-point data otherwise
    const std::size_t n_quads = view.quadratures.size();
    std::vector<std::vector<double>>        table(n_quads), inv_jac(n_quads), jxw(n_quads);
    std::vector<std::vector<std::uint32_t>> cell_row(n_quads);
    std::vector<adaflo_b200_quad_desc>      quads(n_quads);
    for (std::size_t qi = 0; qi < n_quads; ++qi)
      {
        const QuadratureView &qv   = view.quadratures[qi];
        int                   type = 0;
        for (unsigned int b = 0; b < view.n_cell_batches; ++b) type = std::max(type, qv.cell_type(b));
        adaflo_b200_quad_desc &qd = quads[qi];
        qd                        = adaflo_b200_quad_desc{};
        qd.n_q_1d                 = qv.n_q_1d;
        qd.geometry_type          = type;
        if (type == ADAFLO_B200_GEOM_GENERAL)
          {
            const unsigned int n_q = qv.n_q_1d * qv.n_q_1d * qv.n_q_1d;
            inv_jac[qi].resize(static_cast<std::size_t>(n_cells) * 9 * n_q);
            jxw[qi].resize(static_cast<std::size_t>(n_cells) * n_q);
            for (std::uint32_t c = 0; c < n_cells; ++c)
              for (unsigned int q = 0; q < n_q; ++q)
                {
                  double k[9], w;
                  qv.geometry(out.cell_of[c].first, out.cell_of[c].second, q, k, w);
                  for (int r = 0; r < 9; ++r) inv_jac[qi][(static_cast<std::size_t>(c) * 9 + r) * n_q + q] = k[r];
                  jxw[qi][static_cast<std::size_t>(c) * n_q + q] = w;
                }
            qd.inv_jacobian = inv_jac[qi].data();
            qd.jxw          = jxw[qi].data();
          }
        else
          {
            const int                                   row_len = type == ADAFLO_B200_GEOM_CARTESIAN ? 4 : 10;
            std::map<std::vector<double>, std::uint32_t> rows;
            cell_row[qi].resize(n_cells);
            for (std::uint32_t c = 0; c < n_cells; ++c)
              {
                double k[9], det;
                qv.geometry(out.cell_of[c].first, out.cell_of[c].second, 0, k, det);
                std::vector<double> row;
                if (type == ADAFLO_B200_GEOM_CARTESIAN)
                  row = {k[0], k[4], k[8], det};
                else
                  {
                    row.assign(k, k + 9);
                    row.push_back(det);
                  }
                auto it = rows.find(row);
                if (it == rows.end())
                  {
                    it = rows.emplace(row, static_cast<std::uint32_t>(rows.size())).first;
                    table[qi].insert(table[qi].end(), row.begin(), row.end());
                  }
                cell_row[qi][c] = it->second;
              }
            qd.n_table_rows = static_cast<std::uint32_t>(table[qi].size() / row_len);
            qd.table        = table[qi].data();
            qd.cell_row     = cell_row[qi].data();
          }
      }

    adaflo_b200_mf_desc desc = {};
    desc.dim                 = 3;
    desc.n_cells             = n_cells;
    desc.n_spaces            = static_cast<int>(n_spaces);
    desc.spaces              = spaces.data();
    desc.n_quads             = static_cast<int>(n_quads);
    desc.quads               = quads.data();
    out.matrix_free.reset(new MatrixFreeData(ctx, desc));
    // (5) tell the library where its entries live in the host vectors (identity numberings are skipped)
    for (std::size_t s = 0; s < n_spaces; ++s)
      {
        bool identity = true;
        for (std::size_t i = 0; i < out.host_position[s].size() && identity; ++i) identity = out.host_position[s][i] == i;
        if (!identity)
          check(adaflo_b200_mf_set_dof_numbering(out.matrix_free->handle(), static_cast<int>(s), out.host_position[s].data()));
      }
    return out;
  }
} // namespace adaflo_b200
