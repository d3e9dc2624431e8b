//! `Grid` / `SnapResult`: host-side value types with the reference's getters. The flat output order of
//! `generate_grid_sdf` is `get_cell_idx`: z fastest, x slowest.
use crate::{ffi, Point};
#[cfg(feature = "serde")]
use serde::{de::DeserializeOwned, Deserialize, Serialize};

#[derive(Debug, Clone, PartialEq, Eq, PartialOrd, Ord)]
pub enum SnapResult {
    Inside([usize; 3]),
    Outside([usize; 3]),
}

#[derive(Debug, Clone, PartialEq, PartialOrd, Eq, Ord)]
#[cfg_attr(feature = "serde", derive(Serialize, Deserialize))]
#[cfg_attr(feature = "serde", serde(bound = "V: Serialize + DeserializeOwned"))]
pub struct Grid<V: Point> {
    first_cell: V,
    cell_size: V,
    cell_count: [usize; 3],
}

fn arr<V: Point>(v: &V) -> [f32; 3] { [v.x(), v.y(), v.z()] }

impl<V: Point> Grid<V> {
    pub const fn new(first_cell: V, cell_size: V, cell_count: [usize; 3]) -> Self {
        Self { first_cell, cell_size, cell_count }
    }

    /// cell_size = (max - min) / count, first_cell = min + cell_size / 2 — evaluated by the same C helper the
    /// other bindings use, so every language produces bit-identical grids.
    pub fn from_bounding_box(bbox_min: &V, bbox_max: &V, cell_count: [usize; 3]) -> Self {
        let (mn, mx) = (arr(bbox_min), arr(bbox_max));
        let cc = [cell_count[0] as u64, cell_count[1] as u64, cell_count[2] as u64];
        let (mut first, mut size) = ([0f32; 3], [0f32; 3]);
        unsafe { ffi::m2s_grid_from_bounding_box(mn.as_ptr(), mx.as_ptr(), cc.as_ptr(), first.as_mut_ptr(), size.as_mut_ptr()) };
        Self::new(V::new(first[0], first[1], first[2]), V::new(size[0], size[1], size[2]), cell_count)
    }

    pub const fn get_first_cell(&self) -> V { self.first_cell }
    pub const fn get_cell_size(&self) -> V { self.cell_size }
    pub const fn get_cell_count(&self) -> [usize; 3] { self.cell_count }
    pub const fn get_total_cell_count(&self) -> usize { self.cell_count[0] * self.cell_count[1] * self.cell_count[2] }

    pub fn get_last_cell(&self) -> V {
        let (f, s, n) = (arr(&self.first_cell), arr(&self.cell_size), self.cell_count);
        V::new(f[0] + n[0] as f32 * s[0], f[1] + n[1] as f32 * s[1], f[2] + n[2] as f32 * s[2])
    }

    pub fn get_bounding_box(&self) -> (V, V) {
        let (f, s, n) = (arr(&self.first_cell), arr(&self.cell_size), self.cell_count);
        let lo = [f[0] - s[0] * 0.5, f[1] - s[1] * 0.5, f[2] - s[2] * 0.5];
        let hi = [lo[0] + n[0] as f32 * s[0], lo[1] + n[1] as f32 * s[1], lo[2] + n[2] as f32 * s[2]];
        (V::new(lo[0], lo[1], lo[2]), V::new(hi[0], hi[1], hi[2]))
    }

    pub const fn get_cell_idx(&self, cell: &[usize; 3]) -> usize {
        cell[2] + self.cell_count[2] * (cell[1] + self.cell_count[1] * cell[0])
    }

    pub const fn get_cell_integer_coordinates(&self, cell_idx: usize) -> [usize; 3] {
        let plane = self.cell_count[1] * self.cell_count[2];
        [cell_idx / plane, (cell_idx / self.cell_count[2]) % self.cell_count[1], cell_idx % self.cell_count[2]]
    }

    pub fn get_cell_center(&self, cell: &[usize; 3]) -> V {
        let (f, s) = (arr(&self.first_cell), arr(&self.cell_size));
        V::new(f[0] + cell[0] as f32 * s[0], f[1] + cell[1] as f32 * s[1], f[2] + cell[2] as f32 * s[2])
    }

    pub fn snap_point_to_grid(&self, point: &V) -> SnapResult {
        let lo = arr(&self.get_bounding_box().0);
        let (p, s) = (arr(point), arr(&self.cell_size));
        let mut raw = [0isize; 3];
        let mut res = [0usize; 3];
        for i in 0..3 {
            raw[i] = ((p[i] - lo[i]) / s[i]).floor() as isize; // saturating, NaN -> 0
            res[i] = raw[i].clamp(0, self.cell_count[i] as isize - 1) as usize;
        }
        if (0..3).all(|i| raw[i] == res[i] as isize) { SnapResult::Inside(res) } else { SnapResult::Outside(res) }
    }
}
