# R side of the drop-in (not run in this repository's image: no R here).
# In the reference package this replaces lines 50-109 of R/main.r inside res_nmtf_inner(); every other line
# of R/main.r, R/utils.r, R/obtain_bicl.r and R/stability_analysis.r stays as it is.  NAMESPACE gains
#   useDynLib(resnmtf, .registration = TRUE)
# and DESCRIPTION gains  SystemRequirements: CUDA (sm_100a), libresnmtf_b200.

# shared names (the hash objects of produce_indices, R/utils.r:560-601) -> 1-based index pairs
resnmtf_index_maps <- function(indices, names_list) {
  if (is.null(indices)) {
    return(NULL) # the k-extension loop's column_indices = NULL (R/main.r:312)
  }
  n_v <- length(names_list)
  lapply(seq_len(n_v), function(v) {
    lapply(seq_len(n_v), function(w) {
      if (w == v) {
        return(NULL)
      }
      shared <- indices[[v]][[as.character(w)]]
      if (is.null(shared)) {
        return(NULL)
      }
      if (any(is.na(shared))) {
        return(integer(0)) # NA: the pair shares nothing (skipped, R/utils.r:70)
      }
      rbind(match(shared, names_list[[v]]), match(shared, names_list[[w]]))
    })
  })
}

# body of res_nmtf_inner() between init_mats() and obtain_biclusters()
resnmtf_device_loop <- function(data, current_f, current_s, current_g, current_lam, current_mu,
                                phi, xi, psi, row_indices, column_indices, n_iters) {
  row_maps <- resnmtf_index_maps(row_indices, lapply(data, rownames))
  col_maps <- resnmtf_index_maps(column_indices, lapply(data, colnames))
  out <- .Call(
    C_resnmtf_fit, data, current_f, current_s, current_g, current_lam, current_mu,
    phi, xi, psi, row_maps, col_maps,
    if (is.null(n_iters)) NA_integer_ else as.integer(n_iters)
  )
  for (v in seq_along(data)) { # dimnames the reference carries (R/update_steps.r:61-64)
    rownames(out$Fn[[v]]) <- rownames(data[[v]])
    rownames(out$Gn[[v]]) <- colnames(data[[v]])
  }
  list(
    current_f = out$Fn, current_s = out$Sn, current_g = out$Gn, # after normalisation_check()
    current_lam = out$lam, current_mu = out$mu, total_err = out$total_err
  )
}

# jsd_calc() (R/utils.r:95-106) for many column pairs at once: `cols` holds the factor columns, `pairs` is a two-column
# integer matrix of (1-based) column indices.  calculate_f_shuffle_jsd() / check_biclusters() (R/obtain_bicl.r:55-68,
# 113-133) keep their loops but collect the pairs and call this once per view.
resnmtf_jsd_pairs <- function(cols, pairs) {
  storage.mode(cols) <- "double"
  .Call(
    C_resnmtf_jsd_pairs, cols, apply(cols, 2, stats::bw.nrd0), apply(cols, 2, max),
    as.integer(pairs[, 1]), as.integer(pairs[, 2])
  )
}

# The independent fits of one apply_resnmtf() call -- the k sweep (R/main.r:270-299), the shuffled refits of
# obtain_shuffled_f() (R/obtain_bicl.r:31-42), the resamples of stability_check() (R/stability_analysis.r:302-338) -- as
# ONE .Call: the library uploads `data` once and runs the units on one native worker thread per GPU
# (resnmtf_batch_run).  Every random input is drawn HERE, in the reference's order (quirk Q14): the k x k noise of
# init_mats_inner() per view and fit, the sub-sample indices, and one seed per shuffled refit.
#   units: list of lists with elements k, noise | init_f/init_s/init_g, rows/cols, shuffle_seed, renormalise,
#          phi/xi/psi, row_indices/column_indices (the hash objects; converted to index maps here), n_iters
resnmtf_device_batch <- function(data, units, prep = FALSE, n_gpus = 0L) {
  rn <- lapply(data, rownames)
  cn <- lapply(data, colnames)
  units <- lapply(units, function(u) {
    sub_r <- if (is.null(u$rows)) rn else Map(function(nm, i) nm[i], rn, u$rows)
    sub_c <- if (is.null(u$cols)) cn else Map(function(nm, i) nm[i], cn, u$cols)
    u$row_maps <- resnmtf_index_maps(u$row_indices, sub_r)
    u$col_maps <- resnmtf_index_maps(u$column_indices, sub_c)
    u$row_indices <- u$column_indices <- NULL
    u$k <- as.integer(u$k)
    if (!is.null(u$rows)) u$rows <- lapply(u$rows, as.integer)
    if (!is.null(u$cols)) u$cols <- lapply(u$cols, as.integer)
    u$n_iters <- if (is.null(u$n_iters)) NA_integer_ else as.integer(u$n_iters)
    u
  })
  out <- .Call(C_resnmtf_batch, data, isTRUE(prep), units, as.integer(n_gpus))
  for (i in seq_along(out)) { # dimnames the reference carries (R/update_steps.r:61-64)
    u <- units[[i]]
    for (v in seq_along(data)) {
      rownames(out[[i]]$F[[v]]) <- if (is.null(u$rows)) rn[[v]] else rn[[v]][u$rows[[v]]]
      rownames(out[[i]]$G[[v]]) <- if (is.null(u$cols)) cn[[v]] else cn[[v]][u$cols[[v]]]
    }
  }
  out
}

# init_mats_inner() (R/update_steps.r:78-125) with the three pieces of svd(x) it uses computed on the GPU: replace
#   ss <- svd(x[[i]])   by   ss <- resnmtf_svd_topk(x[[i]], k_vec[i])   (ss$u, ss$d, ss$v already truncated to k and made
# non-negative by abs(), which is all the function does with them).
resnmtf_svd_topk <- function(x, k) {
  storage.mode(x) <- "double"
  .Call(C_resnmtf_svd_topk, x, as.integer(k))
}

# bisilhouette::bisilhouette(x, row_clusters, col_clusters, method = distance)$bisil as obtain_biclusters() calls it
# (R/obtain_bicl.r:190-199), distance blocks on the GPU.
resnmtf_bisil <- function(x, row_clusters, col_clusters, method = "euclidean") {
  storage.mode(x) <- storage.mode(row_clusters) <- storage.mode(col_clusters) <- "double"
  m <- match(method, c("euclidean", "manhattan", "cosine")) - 1L
  .Call(C_resnmtf_bisil, x, row_clusters, col_clusters, m)$bisil
}
