#!/usr/bin/env Rscript
# Step 2 of the R pinning kit (never run in this repository's image: there is no R here or on the GPU box --
# profiles/r02_r_probe.txt).  Runs the REFERENCE's own functions on the exported golden inputs and writes what the
# oracle (oracle/resnmtf_oracle.py) and the CUDA path are compared with:
#   - ten sweeps of update_matrices() + calculate_error() (R/update_steps.r:272-319, R/utils.r:157-166) from the explicit
#     initial factors (init_mats() with init_f/s/g given, R/update_steps.r:49-60): per-view error of every sweep and the
#     raw F, S, G, lambda, mu after the tenth -- the loop body of R/main.r:83-108, restated only to keep the per-sweep values;
#   - res_nmtf_inner(..., n_iters = NULL, no_clusts = TRUE) (R/main.r:32-120): the converged, normalised factors;
#   - the convergence loop of R/main.r:51-81 for All_Error (the exported function only returns it together with the
#     bisilhouette, which needs the third-party package);
#   - when `bisilhouette` is installed: res_nmtf_inner(..., spurious = FALSE) -> row_clusters, col_clusters, bisil;
#   - base R's stats::bw.nrd0 / stats::density on the converged factor columns, and jsd_calc() (R/utils.r:95-106) when
#     `philentropy` is installed.
# usage: Rscript make_golden.R <checkout of eso28599/resnmtf> <kit input dir> <output dir>
args <- commandArgs(trailingOnly = TRUE)
stopifnot(length(args) == 3)
ref <- args[1]
in_root <- args[2]
out_root <- args[3]
for (f in c("utils.r", "update_steps.r", "obtain_bicl.r", "stability_analysis.r", "main.r")) {
  source(file.path(ref, "R", f))
}
n_sweeps <- 10

read_manifest <- function(dir) {
  m <- read.table(file.path(dir, "manifest.txt"), stringsAsFactors = FALSE)
  stats::setNames(lapply(seq_len(nrow(m)), function(i) c(m[i, 2], m[i, 3])), m[, 1])
}
read_mat <- function(dir, man, name) {
  d <- man[[name]]
  matrix(readBin(file.path(dir, paste0(name, ".f64")), "double", n = d[1] * d[2], size = 8, endian = "little"),
    nrow = d[1], ncol = d[2]
  )
}
writer <- function(dir) {
  dir.create(dir, recursive = TRUE, showWarnings = FALSE)
  lines <- character(0)
  list(
    put = function(name, x) {
      x <- as.matrix(x)
      storage.mode(x) <- "double"
      writeBin(as.vector(x), file.path(dir, paste0(name, ".f64")), size = 8, endian = "little")
      lines <<- c(lines, paste(name, nrow(x), ncol(x)))
    },
    close = function() writeLines(lines, file.path(dir, "manifest.txt"))
  )
}
# shared-name indices: the reference's reorder_data() (R/utils.r:619-662) needs `rje` and `hash`; without them the same
# structure as plain named lists (star_prod_relevant only does indices[[as.character(w)]], R/utils.r:69)
shared_indices <- function(data, n_v, rn, cn) {
  if (requireNamespace("rje", quietly = TRUE) && requireNamespace("hash", quietly = TRUE)) {
    r <- reorder_data(data, n_v, rn, cn)
    return(list(rows = r$row_indices, cols = r$col_indices, how = "reorder_data"))
  }
  one <- function(names_list) {
    lapply(seq_len(n_v), function(v) {
      out <- list()
      for (w in seq_len(n_v)[-v]) {
        common <- intersect(names_list[[v]], names_list[[w]])
        out[[as.character(w)]] <- if (length(common)) common else NA
      }
      out
    })
  }
  list(rows = one(rn), cols = one(cn), how = "plain lists (rje/hash not installed)")
}

for (case in list.dirs(in_root, recursive = FALSE, full.names = FALSE)) {
  dir <- file.path(in_root, case)
  man <- read_manifest(dir)
  n_v <- as.integer(read_mat(dir, man, "n_views")[1, 1])
  k_vec <- as.integer(read_mat(dir, man, "k")[, 1])
  data <- init_f <- init_s <- init_g <- rn <- cn <- vector("list", n_v)
  for (v in seq_len(n_v)) {
    i <- v - 1
    rn[[v]] <- readLines(file.path(dir, paste0("rn", i, ".txt")))
    cn[[v]] <- readLines(file.path(dir, paste0("cn", i, ".txt")))
    data[[v]] <- read_mat(dir, man, paste0("x", i))
    dimnames(data[[v]]) <- list(rn[[v]], cn[[v]])
    init_f[[v]] <- read_mat(dir, man, paste0("f0_", i))
    init_s[[v]] <- read_mat(dir, man, paste0("s0_", i))
    init_g[[v]] <- read_mat(dir, man, paste0("g0_", i))
  }
  phi <- read_mat(dir, man, "phi") # already symmetrised (init_rest_mats), as res_nmtf_inner receives them
  xi <- read_mat(dir, man, "xi")
  psi <- read_mat(dir, man, "psi")
  idx <- shared_indices(data, n_v, rn, cn)
  w <- writer(file.path(out_root, case))

  # ---- ten sweeps, per-sweep values kept (loop body of R/main.r:83-108) ----
  st <- init_mats(data, n_v, k_vec, init_f, init_g, init_s)
  data_norms <- sapply(data, function(x) norm(x, "F")**2)
  sweep_err <- matrix(0, n_sweeps, n_v)
  for (t in seq_len(n_sweeps)) {
    np <- update_matrices(
      x = data, input_f = st$current_f, input_s = st$current_s, input_g = st$current_g,
      lambda = st$current_lam, mu = st$current_mu, phi = phi, xi = xi, psi = psi,
      row_indices = idx$rows, column_indices = idx$cols
    )
    st <- list(
      current_f = np$output_f, current_s = np$output_s, current_g = np$output_g,
      current_lam = np$output_lam, current_mu = np$output_mu
    )
    sweep_err[t, ] <- calculate_error(data, st$current_f, st$current_s, st$current_g, n_v, data_norms)
  }
  w$put("sweep_errors", sweep_err)
  for (v in seq_len(n_v)) {
    i <- v - 1
    w$put(paste0("f", n_sweeps, "_", i), st$current_f[[v]])
    w$put(paste0("s", n_sweeps, "_", i), st$current_s[[v]])
    w$put(paste0("g", n_sweeps, "_", i), st$current_g[[v]])
    w$put(paste0("lam", n_sweeps, "_", i), st$current_lam[[v]])
    w$put(paste0("mu", n_sweeps, "_", i), st$current_mu[[v]])
  }

  # ---- converged run through the exported function ----
  conv <- res_nmtf_inner(data, idx$rows, idx$cols, init_f, init_s, init_g, k_vec, phi, xi, psi,
    n_iters = NULL, no_clusts = TRUE
  )
  for (v in seq_len(n_v)) {
    i <- v - 1
    w$put(paste0("of_", i), conv$output_f[[v]])
    w$put(paste0("os_", i), conv$output_s[[v]])
    w$put(paste0("og_", i), conv$output_g[[v]])
  }
  # All_Error: the loop of R/main.r:51-81
  st <- init_mats(data, n_v, k_vec, init_f, init_g, init_s)
  total_err <- c()
  err_diff <- 1
  err_temp <- 0
  while (err_diff > 1.0e-6) {
    np <- update_matrices(
      x = data, input_f = st$current_f, input_s = st$current_s, input_g = st$current_g,
      lambda = st$current_lam, mu = st$current_mu, phi = phi, xi = xi, psi = psi,
      row_indices = idx$rows, column_indices = idx$cols
    )
    st <- list(
      current_f = np$output_f, current_s = np$output_s, current_g = np$output_g,
      current_lam = np$output_lam, current_mu = np$output_mu
    )
    mean_err <- mean(calculate_error(data, st$current_f, st$current_s, st$current_g, n_v, data_norms))
    total_err <- c(total_err, mean_err)
    err_diff <- abs(mean_err - err_temp)
    err_temp <- utils::tail(total_err, n = 1)
  }
  w$put("all_error", total_err)

  # ---- binary bicluster matrices and the bisilhouette (third-party package) ----
  if (requireNamespace("bisilhouette", quietly = TRUE)) {
    full <- res_nmtf_inner(data, idx$rows, idx$cols, init_f, init_s, init_g, k_vec, phi, xi, psi,
      n_iters = NULL, spurious = FALSE, no_clusts = FALSE
    )
    w$put("bisil", full$bisil)
    for (v in seq_len(n_v)) {
      w$put(paste0("rows_", v - 1), full$row_clusters[[v]])
      w$put(paste0("cols_", v - 1), full$col_clusters[[v]])
    }
  }
  # ---- KDE / JSD of the spurious-bicluster test on the converged factor columns of view 1 ----
  fcols <- conv$output_f[[1]]
  kk <- ncol(fcols)
  w$put("kde_bw", apply(fcols, 2, stats::bw.nrd0))
  mx <- max(fcols[, 1], fcols[, 2])
  w$put("kde_y_col1", stats::density(fcols[, 1], from = 0, to = mx)$y)
  w$put("kde_y_col2", stats::density(fcols[, 2], from = 0, to = mx)$y)
  if (requireNamespace("philentropy", quietly = TRUE)) {
    jsd <- matrix(0, kk, kk)
    for (a in seq_len(kk)) for (b in seq_len(kk)) if (a != b) jsd[a, b] <- jsd_calc(fcols[, a], fcols[, b])
    w$put("jsd", jsd)
  }
  w$close()
  cat(case, ": sweeps to converge", length(total_err), "; indices via", idx$how, "\n")
}
cat("R", R.version.string, "; BLAS", extSoftVersion()[["BLAS"]], "\n")
