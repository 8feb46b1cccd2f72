"""Vector / Matrix API with factories (dense, sparse, CSC) over torch tensors.

Reference: common/.../math/linalg/{Vector.java:27-170, Matrix.java:27-272, VectorFactory.java:34-72,
MatrixFactory.java:35-85} and the Breeze-backed implementations (breeze/DenseVector.java:288-307
axpy/dot -> netlib BLAS).  The apps of this framework operate on whole tables through fused
kernels; this module keeps the small linear-algebra surface for user code that was written
against the reference API (add/sub/scale/div in place or not, axpy, dot, slice, mmul, transpose,
concat / horzcat, sparse vectors and CSC matrices).  Works on any torch device.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import torch

Scalar = Union[int, float]


class Vector:
    """Dense / sparse float vector API (length, get/set, slice, add/sub/scale/div +-in-place, axpy, dot) over torch
    tensors.  Reference: common/src/main/java/edu/snu/cay/common/math/linalg/Vector.java,
    breeze/DenseVector.java:288-307."""

    def __init__(self, data: torch.Tensor) -> None:
        self.t = data

    # ---- structure ----
    def length(self) -> int:
        return int(self.t.shape[0])

    def is_dense(self) -> bool:
        return not self.t.is_sparse

    def active_size(self) -> int:
        return int(self.t._nnz()) if self.t.is_sparse else self.length()

    def get(self, i: int) -> float:
        return float(self.to_dense().t[i])

    def set(self, i: int, v: Scalar) -> None:
        assert self.is_dense(), "set() on sparse vectors is not supported"
        self.t[i] = v

    def copy(self) -> "Vector":
        return Vector(self.t.clone())

    def to_dense(self) -> "Vector":
        return Vector(self.t.to_dense()) if self.t.is_sparse else self

    def slice(self, start: int, end: int) -> "Vector":
        return Vector(self.to_dense().t[start:end].clone())

    # ---- arithmetic ----
    def add(self, o: Union["Vector", Scalar]) -> "Vector":
        return Vector(self.to_dense().t + (o.to_dense().t if isinstance(o, Vector) else o))

    def addi(self, o: Union["Vector", Scalar]) -> "Vector":
        self.t += o.to_dense().t if isinstance(o, Vector) else o
        return self

    def sub(self, o: Union["Vector", Scalar]) -> "Vector":
        return Vector(self.to_dense().t - (o.to_dense().t if isinstance(o, Vector) else o))

    def subi(self, o: Union["Vector", Scalar]) -> "Vector":
        self.t -= o.to_dense().t if isinstance(o, Vector) else o
        return self

    def scale(self, a: Scalar) -> "Vector":
        return Vector(self.t * a)

    def scalei(self, a: Scalar) -> "Vector":
        self.t *= a
        return self

    def div(self, a: Scalar) -> "Vector":
        return Vector(self.t / a)

    def divi(self, a: Scalar) -> "Vector":
        self.t /= a
        return self

    def axpy(self, a: Scalar, x: "Vector") -> "Vector":
        """In place: this += a * x (x may be sparse)."""
        if x.t.is_sparse:
            xc = x.t.coalesce()
            self.t.index_add_(0, xc.indices()[0], xc.values() * a)
        else:
            self.t.add_(x.t, alpha=a)
        return self

    def dot(self, o: "Vector") -> float:
        if self.t.is_sparse and not o.t.is_sparse:
            return o.dot(self)
        if o.t.is_sparse:
            oc = o.t.coalesce()
            return float((self.to_dense().t[oc.indices()[0]] * oc.values()).sum())
        return float(torch.dot(self.t, o.t))


class Matrix:
    """Dense / CSC float matrix API (rows / columns, transpose, element-wise ops, mmul).  Reference:
    common/src/main/java/edu/snu/cay/common/math/linalg/Matrix.java, breeze/DenseMatrix.java."""

    def __init__(self, data: torch.Tensor) -> None:
        self.t = data

    def rows(self) -> int:
        return int(self.t.shape[0])

    def columns(self) -> int:
        return int(self.t.shape[1])

    def is_dense(self) -> bool:
        return self.t.layout == torch.strided

    def to_dense(self) -> "Matrix":
        return self if self.is_dense() else Matrix(self.t.to_dense())

    def get(self, r: int, c: int) -> float:
        return float(self.to_dense().t[r, c])

    def set(self, r: int, c: int, v: Scalar) -> None:
        self.t[r, c] = v

    def slice_row(self, r: int) -> Vector:
        return Vector(self.to_dense().t[r].clone())

    def slice_column(self, c: int) -> Vector:
        return Vector(self.to_dense().t[:, c].clone())

    def slice_rows(self, a: int, b: int) -> "Matrix":
        return Matrix(self.to_dense().t[a:b].clone())

    def slice_columns(self, a: int, b: int) -> "Matrix":
        return Matrix(self.to_dense().t[:, a:b].clone())

    def put_row(self, r: int, v: Vector) -> None:
        self.t[r] = v.to_dense().t

    def put_column(self, c: int, v: Vector) -> None:
        self.t[:, c] = v.to_dense().t

    def transpose(self) -> "Matrix":
        return Matrix(self.to_dense().t.t().contiguous())

    def copy(self) -> "Matrix":
        return Matrix(self.t.clone())

    def add(self, o: Union["Matrix", Scalar]) -> "Matrix":
        return Matrix(self.to_dense().t + (o.to_dense().t if isinstance(o, Matrix) else o))

    def addi(self, o: Union["Matrix", Scalar]) -> "Matrix":
        self.t += o.to_dense().t if isinstance(o, Matrix) else o
        return self

    def sub(self, o: Union["Matrix", Scalar]) -> "Matrix":
        return Matrix(self.to_dense().t - (o.to_dense().t if isinstance(o, Matrix) else o))

    def subi(self, o: Union["Matrix", Scalar]) -> "Matrix":
        self.t -= o.to_dense().t if isinstance(o, Matrix) else o
        return self

    def mul(self, o: Union["Matrix", Scalar]) -> "Matrix":
        """Element-wise product."""
        return Matrix(self.to_dense().t * (o.to_dense().t if isinstance(o, Matrix) else o))

    def muli(self, o: Union["Matrix", Scalar]) -> "Matrix":
        self.t *= o.to_dense().t if isinstance(o, Matrix) else o
        return self

    def div(self, o: Union["Matrix", Scalar]) -> "Matrix":
        return Matrix(self.to_dense().t / (o.to_dense().t if isinstance(o, Matrix) else o))

    def divi(self, o: Union["Matrix", Scalar]) -> "Matrix":
        self.t /= o.to_dense().t if isinstance(o, Matrix) else o
        return self

    def mmul(self, o: Union["Matrix", Vector]) -> Union["Matrix", Vector]:
        """Matrix-matrix or matrix-vector product (sparse CSC/COO left operands supported)."""
        if isinstance(o, Vector):
            x = o.to_dense().t
            y = torch.sparse.mm(self.t, x[:, None])[:, 0] if not self.is_dense() else self.t @ x
            return Vector(y)
        r = torch.sparse.mm(self.t, o.to_dense().t) if not self.is_dense() else self.t @ o.to_dense().t
        return Matrix(r)


class VectorFactory:
    """Dense-zeros / dense-from-array / sparse constructors.  Reference:
    common/src/main/java/edu/snu/cay/common/math/linalg/VectorFactory.java."""

    def __init__(self, device: str = "cpu", dtype=torch.float32) -> None:
        self.device, self.dtype = device, dtype

    def create_dense_zeros(self, n: int) -> Vector:
        return Vector(torch.zeros(n, dtype=self.dtype, device=self.device))

    def create_dense_ones(self, n: int) -> Vector:
        return Vector(torch.ones(n, dtype=self.dtype, device=self.device))

    def create_dense(self, data: Sequence[float]) -> Vector:
        return Vector(torch.as_tensor(data, dtype=self.dtype, device=self.device).clone())

    def create_sparse_zeros(self, n: int) -> Vector:
        return self.create_sparse([], [], n)

    def create_sparse(self, index: Sequence[int], data: Sequence[float], n: int) -> Vector:
        idx = torch.as_tensor(index, dtype=torch.int64, device=self.device).reshape(1, -1)
        val = torch.as_tensor(data, dtype=self.dtype, device=self.device)
        return Vector(torch.sparse_coo_tensor(idx, val, (n,)).coalesce())

    def concat_dense(self, vectors: List[Vector]) -> Vector:
        return Vector(torch.cat([v.to_dense().t for v in vectors]))


class MatrixFactory:
    """Dense / CSC matrix constructors.  Reference:
    common/src/main/java/edu/snu/cay/common/math/linalg/MatrixFactory.java."""

    def __init__(self, device: str = "cpu", dtype=torch.float32) -> None:
        self.device, self.dtype = device, dtype

    def create_dense_zeros(self, r: int, c: int) -> Matrix:
        return Matrix(torch.zeros((r, c), dtype=self.dtype, device=self.device))

    def create_dense_ones(self, r: int, c: int) -> Matrix:
        return Matrix(torch.ones((r, c), dtype=self.dtype, device=self.device))

    def create_dense(self, data: Sequence[float], r: int, c: int) -> Matrix:
        """Column-major data like the reference (Breeze)."""
        return Matrix(torch.as_tensor(data, dtype=self.dtype, device=self.device).reshape(c, r).t().contiguous())

    def create_csc_zeros(self, r: int, c: int, initial_nonzeros: int = 0) -> Matrix:
        return Matrix(torch.sparse_coo_tensor(torch.zeros((2, 0), dtype=torch.int64, device=self.device),
                                              torch.zeros(0, dtype=self.dtype, device=self.device), (r, c)))

    def horzcat_vec_dense(self, vectors: List[Vector]) -> Matrix:
        return Matrix(torch.stack([v.to_dense().t for v in vectors], dim=1))

    def horzcat_vec_sparse(self, vectors: List[Vector]) -> Matrix:
        """Columns are sparse vectors -> sparse matrix [len, n_vectors]."""
        rows, cols, vals = [], [], []
        for j, v in enumerate(vectors):
            vc = v.t.coalesce() if v.t.is_sparse else v.t.to_sparse().coalesce()
            rows.append(vc.indices()[0])
            cols.append(torch.full_like(vc.indices()[0], j))
            vals.append(vc.values())
        n = vectors[0].length() if vectors else 0
        idx = torch.stack([torch.cat(rows), torch.cat(cols)]) if vectors else torch.zeros((2, 0), dtype=torch.int64)
        return Matrix(torch.sparse_coo_tensor(idx, torch.cat(vals) if vectors else torch.zeros(0), (n, len(vectors))).coalesce())

    def horzcat_mat_dense(self, mats: List[Matrix]) -> Matrix:
        return Matrix(torch.cat([m.to_dense().t for m in mats], dim=1))
